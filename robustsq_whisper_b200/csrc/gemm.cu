// tsw_gemm: argument validation + dispatch between the tcgen05 kernel (gemm_tc.cu) and the fp32-accumulate SIMT
// kernel (gemm_simt.cu).  See include/tsw.h for the contract.
#include <cstdlib>

#include "gemm_common.cuh"

using namespace tsw;

extern "C" size_t tsw_gemm_workspace_bytes(const tsw_gemm_desc* d) { (void)d; return 0; }

extern "C" int tsw_gemm(const tsw_gemm_desc* d, void* workspace, size_t workspace_bytes, tsw_stream_t stream) {
  (void)workspace; (void)workspace_bytes;
  TSW_CHECK_ARG(d != nullptr, "gemm: null descriptor");
  const tsw_gemm_desc& g = *d;
  TSW_CHECK_ARG(g.A && g.B && g.D, "gemm: null operand");
  TSW_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0 && g.batch_outer >= 1 && g.batch_inner >= 1, "gemm: bad sizes M=%lld N=%lld K=%lld", (long long)g.M, (long long)g.N, (long long)g.K);
  TSW_CHECK_ARG((g.a_dtype | 1) == 1 && (g.b_dtype | 1) == 1 && (g.d_dtype | 1) == 1, "gemm: bad dtype");
  TSW_CHECK_ARG(g.kgroups >= 1 || (g.lda >= (g.a_mn_major ? g.M : g.K) && g.ldb >= (g.b_mn_major ? g.N : g.K)), "gemm: leading dimension too small");
  TSW_CHECK_ARG(g.ldd >= g.N, "gemm: leading dimension too small");
  TSW_CHECK_ARG(g.epilogue >= TSW_EPI_NONE && g.epilogue <= TSW_EPI_MUL_AUX, "gemm: bad epilogue");
  TSW_CHECK_ARG((g.epilogue != TSW_EPI_MUL_DGELU && g.epilogue != TSW_EPI_MUL_AUX) || g.aux_in, "gemm: MUL_DGELU / MUL_AUX need aux_in");
  TSW_CHECK_ARG(!g.residual || (g.res_dtype == g.d_dtype && g.ldres >= g.N), "gemm: residual must have the output dtype");
  TSW_CHECK_ARG(g.beta == 0.f || g.beta == 1.f, "gemm: beta must be 0 or 1");
  TSW_CHECK_ARG((g.A2 == nullptr) == (g.B2 == nullptr), "gemm: A2 and B2 go together");
  TSW_CHECK_ARG(!g.A2 || (g.K2 > 0 && g.lda2 >= (g.a_mn_major ? g.M : g.K2) && g.ldb2 >= (g.b_mn_major ? g.N : g.K2)),
                "gemm: second operand pair: bad K2 / leading dimensions");

  const bool grouped = g.kgroups >= 1;
  TSW_CHECK_ARG(g.kgroups >= 0, "gemm: bad kgroups");
  if (grouped) {
    // the leading-dimension check above assumed a dense K: redo it for one group's columns
    const int64_t Kg = g.K / g.kgroups;
    TSW_CHECK_ARG(g.K % g.kgroups == 0 && g.lda >= (g.a_mn_major ? g.M : Kg) && g.ldb >= (g.b_mn_major ? g.N : Kg), "gemm: grouped contraction: bad K / leading dimensions");
  }

  EpiParams ep;
  ep.D = g.D; ep.ldd = g.ldd;
  ep.bias = g.bias;
  ep.residual = g.residual; ep.ldres = g.ldres; ep.res_row_mod = g.res_row_mod;
  ep.aux_in = g.aux_in; ep.aux_out = g.aux_out;
  ep.epilogue = g.epilogue;
  ep.alpha = g.alpha; ep.beta = g.beta; ep.alpha_dev = g.alpha_dev;
  ep.colsum = g.colsum_out;
  ep.M = g.M; ep.N = g.N;
  const int vn = g.d_dtype == TSW_F32 ? 4 : 8;
  auto ok = [&](const void* p, int64_t ld, int64_t so, int64_t si) {
    return !p || (aligned16(p) && ld % vn == 0 && so % vn == 0 && si % vn == 0);
  };
  ep.vec_ok = ok(g.D, g.ldd, g.d_stride_outer, g.d_stride_inner) && ok(g.aux_in, g.ldd, 0, 0) && ok(g.aux_out, g.ldd, 0, 0) &&
              ok(g.residual, g.ldres, g.res_stride_outer, g.res_stride_inner) && (!g.bias || aligned16(g.bias));

  auto ok4 = [&](const void* p, int64_t ld, int64_t so, int64_t si) {
    const uintptr_t al = g.d_dtype == TSW_F32 ? 15u : 7u;
    return !p || ((reinterpret_cast<uintptr_t>(p) & al) == 0 && ld % 4 == 0 && so % 4 == 0 && si % 4 == 0);
  };
  ep.vec4_ok = ok4(g.D, g.ldd, g.d_stride_outer, g.d_stride_inner) && ok4(g.aux_in, g.ldd, 0, 0) && ok4(g.aux_out, g.ldd, 0, 0) &&
               ok4(g.residual, g.ldres, g.res_stride_outer, g.res_stride_inner) && (!g.bias || aligned16(g.bias));

  cudaStream_t st = as_stream(stream);
  if (g.colsum_out && (g.impl == TSW_GEMM_SIMT || !gemm_tc_supported(g, nullptr))) {
    set_error("gemm: colsum_out is a feature of the tcgen05 kernel (bf16 operands that satisfy the TMA constraints)");
    return TSW_E_UNSUPPORTED;
  }
  if (grouped) {
    const char* why = nullptr;
    if (g.impl == TSW_GEMM_SIMT || g.impl == TSW_GEMM_SKINNY || !gemm_tc_supported(g, &why)) {
      set_error("gemm: a grouped contraction runs on the tcgen05 kernel only%s%s", why ? ": " : "", why ? why : "");
      return TSW_E_UNSUPPORTED;
    }
    return gemm_tc_launch(g, ep, st);
  }
  if (g.impl == TSW_GEMM_SIMT) return gemm_simt_launch(g, ep, st);
  if (g.impl == TSW_GEMM_TCGEN05) return gemm_tc_launch(g, ep, st);
  if (g.impl == TSW_GEMM_SKINNY) {
    if (!gemm_skinny_supported(g)) { set_error("gemm(skinny): needs bf16 K-major operands, M <= 32, 256 <= N <= 8192, K %% 8 == 0, a NONE / GELU epilogue"); return TSW_E_UNSUPPORTED; }
    return gemm_skinny_launch(g, ep, st);
  }
  TSW_CHECK_ARG(g.impl == TSW_GEMM_AUTO, "gemm: bad impl %d", g.impl);
  static const bool no_skinny = getenv("TSW_GEMM_NO_SKINNY") != nullptr;
  if (!no_skinny && gemm_skinny_supported(g)) return gemm_skinny_launch(g, ep, st);
  if (gemm_tc_supported(g, nullptr)) return gemm_tc_launch(g, ep, st);
  return gemm_simt_launch(g, ep, st);
}
