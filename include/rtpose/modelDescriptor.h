// ModelDescriptor - same public interface as the reference's include/rtpose/modelDescriptor.h:13-59
// (constructor from a part-name table, a limb sequence and a map index; getters get_number_parts,
// number_limb_sequence, get_limb_sequence, get_map_idx, get_part_name), implemented here from scratch.
// Error behaviour kept: the constructor throws std::runtime_error when limbSequence and mapIdx differ in size
// (modelDescriptor.cpp:30-31) and get_part_name throws std::out_of_range for an unknown index (std::map::at).
#ifndef RTPOSE_MODEL_DESCRIPTOR_H
#define RTPOSE_MODEL_DESCRIPTOR_H
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

class ModelDescriptor {
public:
    ModelDescriptor(const std::map<int, std::string>& partToNameBaseLine, const std::vector<int>& limbSequence,
                    const std::vector<int>& mapIdx)
        : mPartToName(partToNameBaseLine), mLimbSequence(limbSequence), mMapIdx(mapIdx),
          mNumberParts((int)partToNameBaseLine.size() - 1) {
        if (limbSequence.size() != mapIdx.size())
            throw std::runtime_error("limbSequence.size() should be equal to mMapIdx.size()");
        // PAF channels get "<partA>-><partB>(X|Y)" names
        for (size_t l = 0; l + 1 < mLimbSequence.size(); l += 2) {
            const std::string base = mPartToName.at(mLimbSequence[l]) + "->" + mPartToName.at(mLimbSequence[l + 1]);
            mPartToName[mMapIdx[l]] = base + "(X)";
            mPartToName[mMapIdx[l + 1]] = base + "(Y)";
        }
    }
    int get_number_parts() { return mNumberParts; }
    int number_limb_sequence() { return (int)mLimbSequence.size() / 2; }
    const std::vector<int>& get_limb_sequence() { return mLimbSequence; }
    const std::vector<int>& get_map_idx() { return mMapIdx; }
    const std::string& get_part_name(const int partIndex) { return mPartToName.at(partIndex); }

private:
    std::map<int, std::string> mPartToName;
    const std::vector<int> mLimbSequence;
    const std::vector<int> mMapIdx;
    const int mNumberParts;
};
#endif
