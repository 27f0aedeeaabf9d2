// Kernel lookup: one translation unit per LK so the instantiations compile in parallel.
#pragma once
namespace pylda {
// second generation (estep_v2.cuh): W = warps per document group (1, 2, 4, 8); V = 1: 16-warp CTAs
const void* estep_v2_lk1(int J, int W, int V);
const void* estep_v2_lk2(int J, int W, int V);
const void* estep_v2_lk4(int J, int W, int V);
const void* estep_v2_lk8(int J, int W, int V);
const void* estep_v2_lk16(int J, int W, int V);
const void* estep_v2_lk32(int J, int W, int V);
// register-tile generation (estep_rt.cuh); a group holds W * (32/LK) * R rows.  R = NWARPS = 0: the default
// variant (its R is returned in *rows_per_lane); otherwise an explicitly compiled (R, warps per CTA) variant
const void* estep_rt_lk1(int J, int W, int R, int NWARPS, int* rows_per_lane);
const void* estep_rt_lk2(int J, int W, int R, int NWARPS, int* rows_per_lane);
const void* estep_rt_lk4(int J, int W, int R, int NWARPS, int* rows_per_lane);
const void* estep_rt_lk8(int J, int W, int R, int NWARPS, int* rows_per_lane);
const void* estep_rt_lk16(int J, int W, int R, int NWARPS, int* rows_per_lane);
const void* estep_rt_lk32(int J, int W, int R, int NWARPS, int* rows_per_lane);
// second-generation streaming kernel (documents longer than every resident class)
// mode 0: lean, 1: hand-over + chunked staging, 2: hand-over only
const void* estep_stream_lookup(int LK, int J, int mode);
// hybrid register / shared-memory tile kernel for long documents (estep_hy.cuh), clusters of 1, 2, 4, 8 CTAs
const void* estep_hy_lookup(int LK, int J, int* rows_per_lane);
// narrow stages (estep_narrow.cuh): NC = 16 / 8 live columns, G lanes per document
const void* estep_narrow_lookup(int NC, int G, int RPL, int MINB);
// compact stage for long documents (estep_longc.cuh): NC = 32 live columns
const void* estep_longc_lookup(int NC, int ctas_per_sm);
}  // namespace pylda
