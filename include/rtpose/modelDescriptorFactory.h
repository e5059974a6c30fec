// ModelDescriptorFactory - same public interface as the reference's include/rtpose/modelDescriptorFactory.h:10-30
// (enum class Type { MPI_15, COCO_18 }, static createModelDescriptor(type, unique_ptr&)); the tables come from
// libposeengine.so (pe_model_*, include/poseengine.h) so host code and kernels can never disagree.
// Throws std::runtime_error for an unknown type (modelDescriptorFactory.cpp:57-60).
#ifndef RTPOSE_MODEL_DESCRIPTOR_FACTORY_H
#define RTPOSE_MODEL_DESCRIPTOR_FACTORY_H
#include <memory>

#include "../poseengine.h"
#include "modelDescriptor.h"

class ModelDescriptorFactory {
public:
    enum class Type { MPI_15, COCO_18 };
    static void createModelDescriptor(const Type type, std::unique_ptr<ModelDescriptor>& out) {
        if (type != Type::MPI_15 && type != Type::COCO_18) throw std::runtime_error("Undefined ModelDescriptor selected.");
        const int model = type == Type::MPI_15 ? PE_MODEL_MPI_15 : PE_MODEL_COCO_18;
        const int np = pe_model_num_parts(model), nl = pe_model_num_limbs(model);
        std::map<int, std::string> names;
        for (int i = 0; i <= np; i++) names[i] = pe_model_part_name(model, i);
        const int* ls = pe_model_limb_sequence(model);
        const int* mi = pe_model_map_idx(model);
        out.reset(new ModelDescriptor(names, std::vector<int>(ls, ls + 2 * nl), std::vector<int>(mi, mi + 2 * nl)));
    }
};
#endif
