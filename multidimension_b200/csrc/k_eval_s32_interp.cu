// k_eval_s32_interp.cu — instantiations of the fused evaluator with 32-bit value slots (sm_100a).
#include "kernels.cuh"
#include "sigs.hpp"

namespace mdim {

template <class Sig> constexpr const SigInstr* sig_code() { if constexpr (Sig::n > 0) return Sig::code; else return nullptr; }

static const EvalVariant kVariants[] = {
#define X4(Sig, S, V, MAXD, WIDE, MAXR, VPT) \
    {#Sig, (int)sizeof(S), V, MAXD, WIDE ? 1 : 0, MAXR, VPT, sig_code<Sig>(), Sig::n, &k_eval<Sig, S, V, MAXD, WIDE, MAXR, VPT>},
#define X(Sig, S, V, MAXD, WIDE, MAXR) X4(Sig, S, V, MAXD, WIDE, MAXR, 1)
#include "variants_s32_interp.inc"
#undef X
#undef X4
};

const EvalVariant* eval_variants_s32_interp(int* n) {
    *n = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}

}  // namespace mdim
