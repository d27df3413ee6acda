// shapes.cuh -- the expression shapes that get straight-line (StaticShape) instantiations of the hot
// kernels, and the matcher that maps a descriptor to one of them.  Everything else runs through
// the run-time interpreter (DynShape).  The list is the reference's own kernel zoo:
//   GPR/main.py:106-113 (single-input candidates), Multi-Input_GPR/main.py:126-135,521-526 (k1 on the
//   feature dims * k2 on the time dim), the north-star sum kernel and the C1 / C4 benchmark kernels.
// A shape fixes kinds and wiring only; variances, lengthscales, periods and the active-dimension
// masks stay run-time values of the descriptor.
#pragma once
#include <type_traits>

#include "kernel_eval.cuh"

namespace gpb {

#define GPB_L(kind, group) ((kind) | ((group) << 4))
#define GPB_T1(a) (1 | ((a) << 4))
#define GPB_T2(a, b) (2 | ((a) << 4) | ((b) << 8))

constexpr int E = GPB_GROUP_EUCLID, PS = GPB_GROUP_PERIODIC_SQ, DT = GPB_GROUP_DOT;
// id 1..12 (0 = no static shape): all 8 candidates of GPR/main.py:105-114 have one
using Shape_SE = StaticShape<E, -1, GPB_L(GPB_LEAF_SE, 0), -1, -1, -1, GPB_T1(0), 0, 0, 0>;
using Shape_M12 = StaticShape<E, -1, GPB_L(GPB_LEAF_MATERN12, 0), -1, -1, -1, GPB_T1(0), 0, 0, 0>;
using Shape_RQ = StaticShape<E, -1, GPB_L(GPB_LEAF_RQ, 0), -1, -1, -1, GPB_T1(0), 0, 0, 0>;
using Shape_EXP = StaticShape<E, -1, GPB_L(GPB_LEAF_EXPONENTIAL, 0), -1, -1, -1, GPB_T1(0), 0, 0, 0>;
using Shape_SE_M12 = StaticShape<E, -1, GPB_L(GPB_LEAF_SE, 0), GPB_L(GPB_LEAF_MATERN12, 0), -1, -1, GPB_T1(0), GPB_T1(1), 0, 0>;
using Shape_SExM12 = StaticShape<E, -1, GPB_L(GPB_LEAF_SE, 0), GPB_L(GPB_LEAF_MATERN12, 0), -1, -1, GPB_T2(0, 1), 0, 0, 0>;
using Shape_EXPxEXP = StaticShape<E, E, GPB_L(GPB_LEAF_EXPONENTIAL, 0), GPB_L(GPB_LEAF_EXPONENTIAL, 1), -1, -1, GPB_T2(0, 1), 0, 0, 0>;
using Shape_SE_M52_LIN = StaticShape<E, DT, GPB_L(GPB_LEAF_SE, 0), GPB_L(GPB_LEAF_MATERN52, 0), GPB_L(GPB_LEAF_LINEAR, 1), -1,
                                     GPB_T1(0), GPB_T1(1), GPB_T1(2), 0>;
using Shape_SE_M52 = StaticShape<E, -1, GPB_L(GPB_LEAF_SE, 0), GPB_L(GPB_LEAF_MATERN52, 0), -1, -1, GPB_T1(0), GPB_T1(1), 0, 0>;
using Shape_SE_PER = StaticShape<E, PS, GPB_L(GPB_LEAF_SE, 0), GPB_L(GPB_LEAF_SE, 1), -1, -1, GPB_T1(0), GPB_T1(1), 0, 0>;
using Shape_EXP_PER = StaticShape<E, PS, GPB_L(GPB_LEAF_EXPONENTIAL, 0), GPB_L(GPB_LEAF_SE, 1), -1, -1, GPB_T1(0), GPB_T1(1), 0, 0>;
// three groups: Exponential + Periodic(SquaredExponential) + Linear, GPR/main.py:111
using Shape_EXP_PER_LIN = StaticShape<E, PS, GPB_L(GPB_LEAF_EXPONENTIAL, 0), GPB_L(GPB_LEAF_SE, 1), GPB_L(GPB_LEAF_LINEAR, 2), -1,
                                      GPB_T1(0), GPB_T1(1), GPB_T1(2), 0, DT>;

enum ShapeId {
    SHAPE_NONE = 0, SHAPE_SE, SHAPE_M12, SHAPE_RQ, SHAPE_EXP, SHAPE_SE_M12, SHAPE_SExM12, SHAPE_EXPxEXP, SHAPE_SE_M52_LIN,
    SHAPE_SE_M52, SHAPE_SE_PER, SHAPE_EXP_PER, SHAPE_EXP_PER_LIN, SHAPE_COUNT
};

// X(id, policy type, mask of the padded input dimensions DP in {1,2,4,8,16} that get a static instantiation):
// single-input candidates (GPR/main.py) at DP = 1, the multi-input product and the plain SE / Exponential
// at every DP, the north-star / C4 sums at the DPs of those configurations.  Other (shape, DP) pairs run
// through the interpreter.
#define GPB_SHAPE_LIST(X)                                                                              \
    X(SHAPE_SE, Shape_SE, 31) X(SHAPE_M12, Shape_M12, 1) X(SHAPE_RQ, Shape_RQ, 1) X(SHAPE_EXP, Shape_EXP, 31) \
    X(SHAPE_SE_M12, Shape_SE_M12, 1) X(SHAPE_SExM12, Shape_SExM12, 1) X(SHAPE_EXPxEXP, Shape_EXPxEXP, 30)   \
    X(SHAPE_SE_M52_LIN, Shape_SE_M52_LIN, 12) X(SHAPE_SE_M52, Shape_SE_M52, 12) X(SHAPE_SE_PER, Shape_SE_PER, 1) \
    X(SHAPE_EXP_PER, Shape_EXP_PER, 1) X(SHAPE_EXP_PER_LIN, Shape_EXP_PER_LIN, 1)

// switch over the shape id: the body (GPB_SHAPE_BODY_) sees the policy type as SH and the DP mask as
// SH_DPMASK; shapes without a match, and DPs outside the mask, use DynShape (mask 31).
#define GPB_SHAPE_CASE_(ID, T, MASK) case ID: { using SH = T; constexpr unsigned SH_DPMASK = MASK; GPB_SHAPE_BODY_ } break;
#define GPB_DISPATCH_SHAPE(shape)                                                            \
    switch (shape) {                                                                         \
        GPB_SHAPE_LIST(GPB_SHAPE_CASE_)                                                      \
        default: { using SH = DynShape; constexpr unsigned SH_DPMASK = 31u; GPB_SHAPE_BODY_ } break; \
    }
// the policy to instantiate for a given DP inside GPB_SHAPE_BODY_
#define GPB_SH_FOR(DPV) typename std::conditional<(SH_DPMASK & (DPV)) != 0, SH, DynShape>::type

inline int match_shape(const DevKernel& kp) {
    if (shape_matches<Shape_SE>(kp)) return SHAPE_SE;
    if (shape_matches<Shape_M12>(kp)) return SHAPE_M12;
    if (shape_matches<Shape_RQ>(kp)) return SHAPE_RQ;
    if (shape_matches<Shape_EXP>(kp)) return SHAPE_EXP;
    if (shape_matches<Shape_SE_M12>(kp)) return SHAPE_SE_M12;
    if (shape_matches<Shape_SExM12>(kp)) return SHAPE_SExM12;
    if (shape_matches<Shape_EXPxEXP>(kp)) return SHAPE_EXPxEXP;
    if (shape_matches<Shape_SE_M52_LIN>(kp)) return SHAPE_SE_M52_LIN;
    if (shape_matches<Shape_SE_M52>(kp)) return SHAPE_SE_M52;
    if (shape_matches<Shape_SE_PER>(kp)) return SHAPE_SE_PER;
    if (shape_matches<Shape_EXP_PER>(kp)) return SHAPE_EXP_PER;
    if (shape_matches<Shape_EXP_PER_LIN>(kp)) return SHAPE_EXP_PER_LIN;
    return SHAPE_NONE;
}

}  // namespace gpb
