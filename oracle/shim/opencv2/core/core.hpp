#include "../../emvs_cv_shim.h"
