// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Not part of the product.
//
// Prelude that lets line ranges of the reference's own host code compile stand-alone:
//   examples/rtpose/rtpose.cpp:144-152   ColumnCompare
//   examples/rtpose/rtpose.cpp:549-751   connectLimbs        (MPI-15)
//   examples/rtpose/rtpose.cpp:808-1076  connectLimbsCOCO    (COCO-18)
//   src/caffe/util/im2col.cpp:8-56       is_a_ge_zero_and_a_lt_b + im2col_cpu
// together with src/rtpose/modelDescriptor{,Factory}.cpp compiled unmodified.
// It supplies only what those ranges reference from the rest of rtpose.cpp / glog:
// the resolution globals (rtpose.cpp:75-80), the threshold block of `Global` (:106-111),
// MAX_PEOPLE (:88, renderFunctions.h:6) and glog's CHECK/LOG as abort/no-op.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "rtpose/modelDescriptor.h"
#include "rtpose/modelDescriptorFactory.h"

int DISPLAY_RESOLUTION_WIDTH;
int DISPLAY_RESOLUTION_HEIGHT;
int NET_RESOLUTION_WIDTH;
int NET_RESOLUTION_HEIGHT;
const auto MAX_PEOPLE = 96;

struct Global {
    float nms_threshold;
    int connect_min_subset_cnt;
    float connect_min_subset_score;
    float connect_inter_threshold;
    int connect_inter_min_above_threshold;
};
Global global;

struct RefNullStream {
    template <typename T> RefNullStream& operator<<(const T&) { return *this; }
};
#define REF_CHECK_OP(a, op, b) \
    if (!((a)op(b))) { fprintf(stderr, "ref CHECK failed: %s %s %s\n", #a, #op, #b); abort(); } else RefNullStream()
#define CHECK_EQ(a, b) REF_CHECK_OP(a, ==, b)
#define CHECK_GE(a, b) REF_CHECK_OP(a, >=, b)
#define CHECK_LE(a, b) REF_CHECK_OP(a, <=, b)
#define LOG(x) RefNullStream()
