// Host-side interface of the Migrad fit kernels (npswf_migrad.cu, compiled with -fmad=false) for npswf_api.cu.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace npswf {

struct MigradArgs {
    const int *job_list;      // item ids (| N << 27 when list_N == 0)
    const int *job_count;
    int *job_next;            // zeroed cursor
    int list_N;               // pulses of every job of the list, or 0: packed into the entries
    const double *signal, *corr;
    DevCalib cal;
    KParams kp;
    double *wftime, *wfampl, *chi2, *timewf, *amplwf;
    uint8_t *status;
    DeviceCounters *ctr;
};

// workspace classes: 0 -> up to 3 pulses (7 parameters), 1 -> up to 6 (13), 2 -> up to 12 (25)
inline int migrad_class(int N) { return N <= 3 ? 0 : (N <= 6 ? 1 : 2); }
// shared-memory opt-in + resident CTAs per SM of the three instances (current device)
cudaError_t migrad_setup(int occ[3]);
cudaError_t migrad_launch(int cls, int grid, cudaStream_t st, const MigradArgs &a);

}  // namespace npswf
