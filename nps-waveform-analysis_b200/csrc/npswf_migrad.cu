// Translation unit of the Migrad fit kernels.  Built with -fmad=false: the minimiser's decisions must see the chi2
// values and derived quantities an FMA-free x86-64 build of Minuit2 sees (kernel_fit_migrad.cuh).
#include "migrad_launch.hpp"
#include "kernel_fit_migrad.cuh"

namespace npswf {

template <int PMAX>
static cudaError_t setup_one(int *occ)
{
    const size_t smem = sizeof(MgSmem<PMAX>) * MG_WARPS;
    cudaError_t e = cudaFuncSetAttribute(fit_migrad_kernel<PMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, fit_migrad_kernel<PMAX>, MG_THREADS, smem);
}

cudaError_t migrad_setup(int occ[3])
{
    cudaError_t e;
    if ((e = setup_one<7>(&occ[0])) != cudaSuccess) return e;
    if ((e = setup_one<13>(&occ[1])) != cudaSuccess) return e;
    return setup_one<25>(&occ[2]);
}

template <int PMAX>
static void launch_one(int grid, cudaStream_t st, const MigradArgs &a)
{
    fit_migrad_kernel<PMAX><<<grid, MG_THREADS, sizeof(MgSmem<PMAX>) * MG_WARPS, st>>>(
        a.job_list, a.job_count, a.job_next, a.list_N, a.signal, a.corr, a.cal, a.kp, a.wftime, a.wfampl, a.chi2, a.timewf,
        a.amplwf, a.status, a.ctr);
}

cudaError_t migrad_launch(int cls, int grid, cudaStream_t st, const MigradArgs &a)
{
    if (grid <= 0) return cudaSuccess;
    if (cls == 0) launch_one<7>(grid, st, a);
    else if (cls == 1) launch_one<13>(grid, st, a);
    else launch_one<25>(grid, st, a);
    return cudaGetLastError();
}

cudaError_t migrad_thread_setup(int occ[4])
{
    static_assert(MIGRAD_WTAB_ENTRIES == MT_WTAB, "table size");
    cudaError_t e;
    occ[0] = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], fit_migrad_thread_kernel<1>, MT_THREADS, 0)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], fit_migrad_thread_kernel<2>, MT_THREADS, 0)) != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[3], fit_migrad_thread_kernel<3>, MT_THREADS, 0);
}

template <int N>
static void launch_thread(int grid, cudaStream_t st, const MigradArgs &a)
{
    fit_migrad_thread_kernel<N><<<grid, MT_THREADS, 0, st>>>(a.job_list, a.job_count, a.job_next, a.signal, a.corr, a.cal, a.kp,
                                                            a.wftime, a.wfampl, a.chi2, a.timewf, a.amplwf, a.status, a.ctr,
                                                            a.wtab, a.lsb, a.wlow, a.ho_count, a.ho_list);
}

cudaError_t migrad_thread_launch(int N, int grid, cudaStream_t st, const MigradArgs &a)
{
    if (grid <= 0) return cudaSuccess;
    if (N == 1) launch_thread<1>(grid, st, a);
    else if (N == 2) launch_thread<2>(grid, st, a);
    else launch_thread<3>(grid, st, a);
    return cudaGetLastError();
}

int migrad_wtab_floor(double lsb)
{
    const double w0 = mg::inv_err(0.0);
    int c = 0;
    while (c < MT_WTAB && mg::inv_err((double)c * lsb) == w0) c++;
    return c;
}

cudaError_t migrad_build_wtab(double *d_wtab, double lsb, cudaStream_t st)
{
    mg_wtab_kernel<<<(MT_WTAB + 255) / 256, 256, 0, st>>>(d_wtab, MT_WTAB, lsb);
    return cudaGetLastError();
}

}  // namespace npswf
