// TEST SHIM (not image_geometry): the accessors MapperEMVS reads from a PinholeCameraModel (mapper_emvs_stereo.cpp:34-48).
#pragma once
namespace cv {
struct Size {
  int width = 0, height = 0;
};
}  // namespace cv
namespace image_geometry {
class PinholeCameraModel {
 public:
  PinholeCameraModel(int w, int h, double fx, double fy, double cx, double cy) : size_{w, h}, fx_(fx), fy_(fy), cx_(cx), cy_(cy) {}
  cv::Size fullResolution() const { return size_; }
  double fx() const { return fx_; }
  double fy() const { return fy_; }
  double cx() const { return cx_; }
  double cy() const { return cy_; }

 private:
  cv::Size size_;
  double fx_, fy_, cx_, cy_;
};
}  // namespace image_geometry
