// oracle/ref_stub -- TEST INFRASTRUCTURE.  fast_rcnn_layers.hpp declares SmoothL1LossLayer on top
// of LossLayer; it is never instantiated here, the base only has to exist.
#ifndef MNC_REF_STUB_LOSS_LAYER_HPP_
#define MNC_REF_STUB_LOSS_LAYER_HPP_
#include "caffe/layer.hpp"
namespace caffe {
template <typename Dtype>
class LossLayer : public Layer<Dtype> {
 public:
  explicit LossLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
};
}  // namespace caffe
#endif
