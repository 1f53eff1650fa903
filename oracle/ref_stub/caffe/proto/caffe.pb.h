// oracle/ref_stub -- TEST INFRASTRUCTURE.  Stand-in for the protoc-generated caffe.pb.h: only the
// messages the four MNC / Fast R-CNN layers read (caffe-mnc/src/caffe/proto/caffe.proto:
// ROIPoolingParameter, ROIWarpingParameter :1022-1030, MaskResizeParameter :1032-1035), with the
// proto defaults.
#ifndef MNC_REF_STUB_CAFFE_PB_H_
#define MNC_REF_STUB_CAFFE_PB_H_
namespace caffe {
enum Phase { TRAIN = 0, TEST = 1 };
struct ROIPoolingParameter {
  unsigned pooled_h_, pooled_w_; float spatial_scale_;
  ROIPoolingParameter() : pooled_h_(0), pooled_w_(0), spatial_scale_(1.f) {}
  unsigned pooled_h() const { return pooled_h_; }
  unsigned pooled_w() const { return pooled_w_; }
  float spatial_scale() const { return spatial_scale_; }
};
struct ROIWarpingParameter {
  unsigned pooled_h_, pooled_w_; float spatial_scale_;
  ROIWarpingParameter() : pooled_h_(0), pooled_w_(0), spatial_scale_(1.f) {}
  unsigned pooled_h() const { return pooled_h_; }
  unsigned pooled_w() const { return pooled_w_; }
  float spatial_scale() const { return spatial_scale_; }
};
struct MaskResizeParameter {
  unsigned output_height_, output_width_;
  MaskResizeParameter() : output_height_(1), output_width_(1) {}
  unsigned output_height() const { return output_height_; }
  unsigned output_width() const { return output_width_; }
};
struct LayerParameter {
  Phase phase_;
  ROIPoolingParameter roi_pooling_param_;
  ROIWarpingParameter roi_warping_param_;
  MaskResizeParameter mask_resize_param_;
  LayerParameter() : phase_(TEST) {}
  Phase phase() const { return phase_; }
  const ROIPoolingParameter& roi_pooling_param() const { return roi_pooling_param_; }
  const ROIWarpingParameter& roi_warping_param() const { return roi_warping_param_; }
  const MaskResizeParameter& mask_resize_param() const { return mask_resize_param_; }
};
}  // namespace caffe
#endif
