// CTA-pair (cta_group::2) generation of the persistent GRU kernels (fn_gru_tc2.cu), dispatched from fn_gru_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include "../../include/fadernets_b200.h"

// true when the pair kernels serve this launch shape (two 128-row batch tiles, H a multiple of 64, the grid fits)
bool fn_gru2_eligible(bool bwd, int n_chains, int B, int H);
int fn_gru2_run(bool bwd, const FnGruChainBf16* chains, int n_chains, int B, int T, int H, void* barrier_ws, size_t ws_bytes,
                cudaStream_t st);
long long* fn_gru_dbg_ptr();     // device buffer of fn_gru_debug_timeline, or NULL (defined in fn_gru_tc.cu)
