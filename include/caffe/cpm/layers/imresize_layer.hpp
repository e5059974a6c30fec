// caffe::ImResizeLayer<Dtype> accessor shim over the C ABI: the subset of
// include/caffe/cpm/layers/imresize_layer.hpp:20-29 that examples/rtpose/rtpose.cpp uses (:198-202).
#ifndef CAFFE_IMRESIZE_LAYER_HPP_
#define CAFFE_IMRESIZE_LAYER_HPP_
#include "../../../poseengine.h"
namespace caffe {
template <typename Dtype>
class ImResizeLayer {
public:
    explicit ImResizeLayer(pe_engine* engine) : e_(engine) {}
    const char* type() const { return "ImResize"; }
    void SetStartScale(float astart_scale) { pe_resize_set_start_scale(e_, astart_scale); }
    void SetScaleGap(float ascale_gap) { pe_resize_set_scale_gap(e_, ascale_gap); }
    float GetStartScale() { return pe_resize_get_start_scale(e_); }
    float GetScaleGap() { return pe_resize_get_scale_gap(e_); }
private:
    pe_engine* e_;
};
}  // namespace caffe
#endif
