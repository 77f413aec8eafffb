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
    // thread-per-fit kernel only: inverse-error table by |count|, the ADC step, the hand-over list for traces off the lattice
    const double *wtab = nullptr;
    double lsb = 0;
    int wlow = 0;             // table entries below this index all hold the constant-error weight
    int *ho_count = nullptr, *ho_list = nullptr;
};

// workspace classes: 0 -> up to 3 pulses (7 parameters), 1 -> up to 6 (13), 2 -> up to 12 (25)
inline int migrad_class(int N) { return N <= 3 ? 0 : (N <= 6 ? 1 : 2); }
// shared-memory opt-in + resident CTAs per SM of the three instances (current device)
cudaError_t migrad_setup(int occ[3]);
cudaError_t migrad_launch(int cls, int grid, cudaStream_t st, const MigradArgs &a);
// thread-per-fit instances (N = 1, 2, 3 pulses, lattice traces): resident CTAs per SM, launch, and the table they index
constexpr int MIGRAD_WTAB_ENTRIES = 8192;
cudaError_t migrad_thread_setup(int occ[4]);
cudaError_t migrad_thread_launch(int N, int grid, cudaStream_t st, const MigradArgs &a);
cudaError_t migrad_build_wtab(double *d_wtab, double lsb, cudaStream_t st);
int migrad_wtab_floor(double lsb);   // first |count| whose error is above the constant floor of T2:952-954

}  // namespace npswf
