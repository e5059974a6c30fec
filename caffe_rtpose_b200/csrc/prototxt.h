// Deploy-prototxt reader (protobuf text format) and the network definition the execution plan is built from.
// Replaces, for this path, ReadNetParamsFromTextFileOrDie + UpgradeNetAsNeeded + Net::Init of the reference
// (src/caffe/net.cpp:30-50, src/caffe/util/upgrade_proto.cpp:957): new caffe::Net(proto, TEST) at rtpose.cpp:183.
#pragma once
#include <string>
#include <vector>

namespace pe {

struct ProtoLayer {
    std::string name, type;                       // type as in V2 prototxts: "Convolution", "ReLU", "Pooling", "Concat", "ImResize", "Nms"
    std::vector<std::string> bottoms, tops;
    // ConvolutionParameter (caffe.proto: num_output, pad, kernel_size, stride; defaults pad 0, stride 1)
    int num_output = 0, kernel = 0, pad = 0, stride = 1;
    // PoolingParameter (pool default MAX = 0, stride default 1, pad default 0)
    int pool_method = 0;
    // ConcatParameter.axis (default 1)
    int concat_axis = 1;
    // NmsParameter (caffe.proto:1471-1476) and ImResizeParameter (:1478-1484) with their proto defaults
    float nms_threshold = 0.5f; int nms_max_peaks = 20, nms_num_parts = 15;
    float resize_factor = 0.f, resize_start_scale = 1.f, resize_scale_gap = 0.1f;
};
struct NetDef {
    std::string name;
    std::vector<std::string> inputs;
    std::vector<int> input_dims;
    std::vector<ProtoLayer> layers;
};

// 0 on success; err describes the first problem (file, syntax, unsupported field value).
int parse_prototxt_file(const char* path, NetDef& out, std::string& err);
int parse_prototxt_text(const std::string& text, NetDef& out, std::string& err);
// model/{coco,mpi}/pose_deploy_linevec.prototxt rebuilt in code (the graph pe_create uses when no prototxt is given);
// stages = number of CPM stages (6 in the default files; 1, 2, 4 = model/mpi/pose_deploy_linevec_{1,2,4}.prototxt)
NetDef builtin_netdef(int model, int stages);

}  // namespace pe
