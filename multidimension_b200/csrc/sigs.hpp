// sigs.hpp — compile-time op-tree signatures with a pre-instantiated fused kernel.
//
// A signature is the (opcode, dtype, op, aux) sequence of the postfix device program; run-time
// data (pointers, strides, immediates, predicates) still come from the Program parameter.  This is
// the device-side counterpart of the reference monomorphising `Zip<Zip<A,B,Mul>,Scalar,Add>` etc.
// into one specialised collect loop.  Anything not listed runs through the interpreter.
#pragma once
#include "exec.cuh"

namespace mdim {

#define MDIM_SIG(NAME, ...)                                                   \
    struct NAME {                                                             \
        static constexpr SigInstr code[] = {__VA_ARGS__};                     \
        static constexpr int n = (int)(sizeof(code) / sizeof(SigInstr));      \
    };

#define I_LV(dt) {OPC_LEAF_VEC, dt, 0, 0}
#define I_LVC(dt) {OPC_LEAF_VEC, dt, 0, 1} /* re-read operand: L1-allocating loads */
#define I_LB(dt) {OPC_LEAF_BCAST, dt, 0, 0}
#define I_LS(dt) {OPC_LEAF_STRIDED, dt, 0, 0}
#define I_CONST(dt) {OPC_CONST, dt, 0, 0}
#define I_BIN(dt, op) {OPC_BINARY, dt, op, dt}
#define I_SELECT(dt) {OPC_SELECT, dt, 0, 0}
#define I_GATHER1(dt) {OPC_GATHER, dt, 0, 1}
#define I_IOTA(dt) {OPC_IOTA, dt, 0, 0}
#define I_FOLD_BEGIN(dt) {OPC_FOLD_BEGIN, dt, 0, 0}
#define I_FOLD_STEP(dt, op) {OPC_FOLD_STEP, dt, op, dt}

// ---- 32-bit slot programs (f32) ----------------------------------------------------------------
MDIM_SIG(SigCopyF32, I_LV(MDIM_F32))
MDIM_SIG(SigAddF32, I_LV(MDIM_F32), I_LV(MDIM_F32), I_BIN(MDIM_F32, MDIM_ADD))
MDIM_SIG(SigSubF32, I_LV(MDIM_F32), I_LV(MDIM_F32), I_BIN(MDIM_F32, MDIM_SUB))
MDIM_SIG(SigMulF32, I_LV(MDIM_F32), I_LV(MDIM_F32), I_BIN(MDIM_F32, MDIM_MUL))
MDIM_SIG(SigDivF32, I_LV(MDIM_F32), I_LV(MDIM_F32), I_BIN(MDIM_F32, MDIM_DIV))
MDIM_SIG(SigAddCF32, I_LV(MDIM_F32), I_CONST(MDIM_F32), I_BIN(MDIM_F32, MDIM_ADD))
MDIM_SIG(SigSubCF32, I_LV(MDIM_F32), I_CONST(MDIM_F32), I_BIN(MDIM_F32, MDIM_SUB))
MDIM_SIG(SigMulCF32, I_LV(MDIM_F32), I_CONST(MDIM_F32), I_BIN(MDIM_F32, MDIM_MUL))
MDIM_SIG(SigDivCF32, I_LV(MDIM_F32), I_CONST(MDIM_F32), I_BIN(MDIM_F32, MDIM_DIV))
// BASELINE config 2: a.zip(b).map(|(x,y)| x*y+1)  ==  a*b + Scalar(1.0)
MDIM_SIG(SigMulAddCF32, I_LV(MDIM_F32), I_LV(MDIM_F32), I_BIN(MDIM_F32, MDIM_MUL), I_CONST(MDIM_F32), I_BIN(MDIM_F32, MDIM_ADD))
// BASELINE config 4b: a - mean.iso::<(usize,usize,())>()  and  a - (sums / Scalar(256.0)).iso()
MDIM_SIG(SigAddBcastF32, I_LV(MDIM_F32), I_LB(MDIM_F32), I_BIN(MDIM_F32, MDIM_ADD))
MDIM_SIG(SigSubBcastF32, I_LV(MDIM_F32), I_LB(MDIM_F32), I_BIN(MDIM_F32, MDIM_SUB))
MDIM_SIG(SigMulBcastF32, I_LV(MDIM_F32), I_LB(MDIM_F32), I_BIN(MDIM_F32, MDIM_MUL))
MDIM_SIG(SigDivBcastF32, I_LV(MDIM_F32), I_LB(MDIM_F32), I_BIN(MDIM_F32, MDIM_DIV))
// x (op) row-vector broadcast over the outer axes (bias add / column scaling): the vector stays in L1
MDIM_SIG(SigAddVecBcastF32, I_LV(MDIM_F32), I_LVC(MDIM_F32), I_BIN(MDIM_F32, MDIM_ADD))
MDIM_SIG(SigSubVecBcastF32, I_LV(MDIM_F32), I_LVC(MDIM_F32), I_BIN(MDIM_F32, MDIM_SUB))
MDIM_SIG(SigMulVecBcastF32, I_LV(MDIM_F32), I_LVC(MDIM_F32), I_BIN(MDIM_F32, MDIM_MUL))
MDIM_SIG(SigSubBcastDivF32, I_LV(MDIM_F32), I_LB(MDIM_F32), I_CONST(MDIM_F32), I_BIN(MDIM_F32, MDIM_DIV), I_BIN(MDIM_F32, MDIM_SUB))
// BASELINE config 5: transpose -> diagonal(0.0) -> zip(w broadcast) -> x*y+1
MDIM_SIG(SigDiagMulAddCF32, I_LB(MDIM_F32), I_SELECT(MDIM_F32), I_LVC(MDIM_F32), I_BIN(MDIM_F32, MDIM_MUL), I_CONST(MDIM_F32),
         I_BIN(MDIM_F32, MDIM_ADD))
// v.diagonal(zero) of a vector or matrix (inner axis broadcast or strided is interpreted; this is the LB form)
MDIM_SIG(SigDiagF32, I_LB(MDIM_F32), I_SELECT(MDIM_F32))

// rows().map(|r| fold r.each(..)) over any axis that is not a contiguous last axis (that shape has its own
// kernel, k_fold_rows): a fold over the OUTERMOST axis (vector loads, e.g. the per-rank partial of a
// sharded-axis reduction) and over a strided axis.  The loop is a real loop, sequential, in index order.
MDIM_SIG(SigFoldAddVecF32, I_FOLD_BEGIN(MDIM_F32), I_LV(MDIM_F32), I_FOLD_STEP(MDIM_F32, MDIM_ADD))
MDIM_SIG(SigFoldAddStridedF32, I_FOLD_BEGIN(MDIM_F32), I_LS(MDIM_F32), I_FOLD_STEP(MDIM_F32, MDIM_ADD))
MDIM_SIG(SigFoldMulVecF32, I_FOLD_BEGIN(MDIM_F32), I_LV(MDIM_F32), I_FOLD_STEP(MDIM_F32, MDIM_MUL))

// ---- 64-bit slot programs ----------------------------------------------------------------------
MDIM_SIG(SigCopyU64, I_LV(MDIM_U64))
MDIM_SIG(SigIotaU64, I_IOTA(MDIM_U64))
// BASELINE config 3: idx.compose(src): Array<usize,usize> selecting from Array<usize,f32>
MDIM_SIG(SigGatherF32, I_LV(MDIM_U64), I_GATHER1(MDIM_F32))
MDIM_SIG(SigGatherU64, I_LV(MDIM_U64), I_GATHER1(MDIM_U64))
MDIM_SIG(SigCopyF64, I_LV(MDIM_F64))
MDIM_SIG(SigAddF64, I_LV(MDIM_F64), I_LV(MDIM_F64), I_BIN(MDIM_F64, MDIM_ADD))
MDIM_SIG(SigMulF64, I_LV(MDIM_F64), I_LV(MDIM_F64), I_BIN(MDIM_F64, MDIM_MUL))
MDIM_SIG(SigAddU64, I_LV(MDIM_U64), I_LV(MDIM_U64), I_BIN(MDIM_U64, MDIM_ADD))

}  // namespace mdim
