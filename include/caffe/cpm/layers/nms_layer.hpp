// caffe::NmsLayer<Dtype> accessor shim over the C ABI: the subset of include/caffe/cpm/layers/nms_layer.hpp:21-27
// that examples/rtpose/rtpose.cpp uses (GetMaxPeaks :195, GetNumParts :207, SetThreshold :1145).
#ifndef CAFFE_NMS_LAYER_HPP_
#define CAFFE_NMS_LAYER_HPP_
#include "../../../poseengine.h"
namespace caffe {
template <typename Dtype>
class NmsLayer {
public:
    explicit NmsLayer(pe_engine* engine) : e_(engine) {}
    const char* type() const { return "Nms"; }
    int GetMaxPeaks() const { return pe_nms_get_max_peaks(e_); }
    int GetNumParts() const { return pe_nms_get_num_parts(e_); }
    float GetThreshold() const { return pe_nms_get_threshold(e_); }
    void SetThreshold(float threshold) { pe_nms_set_threshold(e_, threshold); }
private:
    pe_engine* e_;
};
}  // namespace caffe
#endif
