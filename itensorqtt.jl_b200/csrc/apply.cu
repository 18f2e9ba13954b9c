// placeholder, replaced below
#include "common.cuh"
namespace qtt {
void apply_mpo_mps(Ctx*, const Mpo&, const Mps&, double, int, int, Mps&) { fail(QTT_ERR_UNSUPPORTED, "qtt_apply not built yet"); }
void mps_inner(Ctx*, const Mps&, const Mps&, double*) { fail(QTT_ERR_UNSUPPORTED, "qtt_inner not built yet"); }
void mps_residual(Ctx*, const Mpo&, const Mps&, const Mps&, double*, double*) { fail(QTT_ERR_UNSUPPORTED, "qtt_residual not built yet"); }
}
