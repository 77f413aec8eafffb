// libnpswf.so — C ABI (include/npswf.h) over the sm_100a kernels.  Host side of the drop-in
// boundary: handle, per-device calibration + workspaces, chunked double-buffered H2D / compute /
// D2H pipeline, event sharding over devices (contiguous ranges, one host thread per device, no
// collective).  There is no CPU fallback: compute entry points fail if CUDA fails.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/npswf.h"
#include "common.cuh"
#include "kernel_front.cuh"
#include "kernel_search.cuh"
#include "kernel_fit.cuh"
#include "kernel_fit_small.cuh"
#include "kernel_fit_thread.cuh"
#include "kernel_fit_vm.cuh"
#include "kernel_aux.cuh"
#include "migrad_launch.hpp"
#include "host_pack.hpp"
#include <chrono>
#include <memory>

using namespace npswf;

namespace {

std::string g_create_error;

struct Workspace {  // per (device, pipeline stage) buffers for up to `cap` events
    int64_t cap = 0;       // events the scratch buffers and job lists hold
    int64_t io_cap = 0;    // events the host-path staging buffers (signal ... status) hold: the host-path chunk limit
    double *signal = nullptr;
    int16_t *counts = nullptr;
    double *packed = nullptr;       // packed hcana stream of a chunk (npswf_analyze_batch_packed)
    long long *poffs = nullptr;     // its event offsets
    size_t packed_cap = 0;
    int32_t *pres = nullptr;
    double *corr = nullptr;
    float *mf = nullptr;
    double *minsig = nullptr;
    uint8_t *flags = nullptr;
    int32_t *wfnpulse = nullptr;
    double *wftime = nullptr, *wfampl = nullptr, *chi2 = nullptr, *timewf = nullptr, *amplwf = nullptr;
    uint8_t *status = nullptr;
    uint8_t *mask = nullptr;
    int *fit_count = nullptr;     // [128]: jobs per multiplicity, [16 + N]: job cursors, [32 + N]: continuation counts, [48 + N]: their cursors, [64 + N] / [80 + N]: second-level hand-over counts / cursors
    int *cont_list = nullptr;     // [3][cap*B] fits handed from fit_thread_kernel to fit_small_kernel (N = 1, 2, 3)
    double *cont_state = nullptr; // [3][cap*B][10] their LM state
    int *bucket_count = nullptr;  // [13][B] jobs per (multiplicity, block)
    int *fit_list = nullptr;      // [13][B][cap] bucketed item ids
    int *fit_dense = nullptr;     // [13][cap*B] block-major dense job lists (what the fit kernels read)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_cmp = nullptr, ev_out = nullptr;  // host-path pipeline: uploaded / computed / downloaded
    // flat (reference-layout) outputs: pulses per event, their prefix sums, the packed pulses; host copy of the sums
    int *ev_total = nullptr, *ev_off = nullptr;
    double *flat_t = nullptr, *flat_a = nullptr;
    int *h_ev_off = nullptr;          // pinned, [cap + 1]
    cudaEvent_t ev_tot = nullptr;     // h_ev_off has arrived
    bool io = false;  // has the signal/pres/output staging buffers (host-buffer API) or only scratch
};

struct DevSlot {
    int device = 0;
    int sm_count = 148;
    DevCalib cal{};
    std::vector<void *> owned;
    Workspace ws[3];   // [0], [1]: both paths; [2]: third stage of the host pipeline (allocated with the host staging buffers)
    DeviceCounters *ctr = nullptr;
    const double *gold1 = nullptr;  // [2][138] first-iteration Gold denominators and reciprocals
    cudaStream_t own_stream = nullptr, copy_in = nullptr, copy_out = nullptr;
    cudaStream_t fit_stream[4] = {nullptr, nullptr, nullptr, nullptr};  // N = 1 | N = 2 | N = 3 | N >= 4 run concurrently
    cudaEvent_t fit_fork = nullptr, fit_join[4] = {nullptr, nullptr, nullptr, nullptr}, chunk_join[2] = {nullptr, nullptr};
    bool fit_concurrent = true;  // env NPSWF_FIT_CONCURRENT=0 serialises the fit kernels on the caller's stream
    bool fit_thread = true;     // thread-per-fit kernels for N = 1, 2 (env NPSWF_FIT_THREAD=0 selects the sub-warp kernels)
    int occ_fit_thread[5] = {0, 2, 2, 2, 2};   // [4]: N = 4..6
    int fit_thread_maxocc = 0;  // env NPSWF_FIT_THREAD_OCC: cap on resident CTAs per SM (fewer CTAs leave more L1)
    int occ_vm_thread[7] = {0, 2, 2, 2, 2, 2, 2};   // fit_vm_thread_kernel<1..6>
    int occ_migrad[3] = {2, 1, 1};   // fit_migrad_kernel<7, 13, 25>
    int occ_migrad_thread[4] = {0, 4, 4, 4};   // fit_migrad_thread_kernel<1, 2, 3>
    double *mg_wtab = nullptr;       // inverse error by |ADC count| (thread-per-fit Migrad kernels)
    double mg_wtab_lsb = 0;          // the ADC step the table was built for
    int mg_wlow = 0;                 // entries below this index hold the constant-error weight
    std::vector<int> local_cpus;     // cores of the NUMA node the GPU hangs off (within this process's affinity mask); empty = unknown
    cudaEvent_t last_use = nullptr;  // end of the last device-path call: later calls order themselves behind it (shared scratch, job lists, fit streams)
    bool last_use_valid = false;
    bool migrad_thread = true;       // env NPSWF_MIGRAD_THREAD=0: every Migrad fit through the warp-per-fit kernel
    int occ_front = 2, occ_search = 4, occ_fit_big = 1, occ_fit_mid = 2, occ_fit_small[4] = {0, 4, 4, 2};  // resident CTAs per SM
    std::vector<cudaEvent_t> prof_events;  // 4 per profiled chunk: start, after front, after search, after fits
    std::vector<cudaEvent_t> prof_pool;
    // lossless int16 transport of the binary64 host layout (host_pack.hpp): pool, three pinned staging buffers
    std::unique_ptr<PackPool> packer;
    int16_t *stage[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev[3] = {nullptr, nullptr, nullptr};   // upload out of the staging buffer has completed
    bool stage_busy[3] = {false, false, false};
    size_t stage_cap = 0;                                    // samples per staging buffer
    int64_t packed_chunks = 0, raw_chunks = 0;
    double pack_seconds = 0, pack_bytes = 0;
    double pack_rate = 0;                                    // running estimate, bytes of doubles per second
    bool pack_unavailable = false;                           // the pinned staging buffers could not be allocated
    unsigned pack_probe = 0;                                 // chunks sent raw because the packer is slow
    // measured rate of the raw (binary64) uploads from pinned memory: CUDA events around the raw part of a chunk
    double dma_rate = 48.0e9;
    cudaEvent_t dma_ev[2] = {nullptr, nullptr};
    bool dma_pending = false;
    double dma_bytes = 0;
};

}  // namespace

struct npswf_handle {
    NpsWfConfig cfg;
    KParams kp;
    std::vector<int> devices;
    std::vector<DevSlot> slots;
    std::vector<double> mfyref, mfint, spline, timeref;
    std::string err;
    NpsWfCounters host_ctr{};
    int64_t chunk = 1184;  // host path: largest chunk (8 x 148 SMs); the pipeline wants 8-12 chunks per call
    int64_t dev_cap = 4736; // device path: largest chunk.  A call is cut into two chunks (>= 1184 events each) up to this size:
                            // the fit kernels' tails cost the same for a short and a long job list (fit stage per 2 368 events:
                            // 8.97 ms in chunks of 1 184, 7.76 in 2 368, 7.03 in 4 736)
    bool chunk_fixed = false;   // cfg.chunk_events given: every path uses exactly that chunk size
    std::mutex mu;
    bool profiling = false;
    int fit_mode = NPSWF_FIT_FAST;
    bool unit_knots = true;            // interpX of every block is 0..109: floor() indexing in the fast kernels
    int pack_mode = 1;                 // 0 off, 1 auto (on while it is faster than the raw upload), 2 always
    bool chunk_ramp = true;            // env NPSWF_CHUNK_RAMP=0: equal chunks in the host pipeline
    int pack_threads = 0;              // host threads per device
    double pack_lsb = 1000.0 / 4096;   // ADCtomV, T2:357
    double stage_ms[3] = {0, 0, 0};  // front, search, fit
    int64_t stage_chunks = 0;
};

namespace {

// Cores of the NUMA node a device is attached to, from sysfs (PCI bus id -> numa_node -> cpulist), intersected with
// the affinity mask the process was started with.  The packer threads and the pinned staging buffers of a device are
// placed there: with several GPUs on a two-socket host a staging buffer on the far socket halves the upload rate.
// Empty when the topology is not visible (single node, container without sysfs) or NPSWF_NUMA_BIND=0.
std::vector<int> device_local_cpus(int device)
{
    std::vector<int> out;
    if (getenv("NPSWF_NUMA_BIND") && atoi(getenv("NPSWF_NUMA_BIND")) == 0) return out;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) { (void)cudaGetLastError(); return out; }
    for (char *c = bus; *c; c++) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    int node = -1;
    if (f) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    if (node < 0) return out;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return out;
    char buf[1024] = {0};
    if (!fgets(buf, sizeof buf, f)) buf[0] = 0;
    fclose(f);
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return out;
    for (char *tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int n = sscanf(tok, "%d-%d", &a, &b);
        if (n == 1) b = a;
        if (n >= 1)
            for (int c = a; c <= b && c < CPU_SETSIZE; c++)
                if (CPU_ISSET(c, &allowed)) out.push_back(c);
    }
    return out;
}

// scope guard: the calling thread runs next to the device while it packs and allocates staging memory
struct ThreadBind {
    cpu_set_t saved;
    bool active = false;
    explicit ThreadBind(const std::vector<int> &cpus)
    {
        if (cpus.empty()) return;
        if (pthread_getaffinity_np(pthread_self(), sizeof saved, &saved) != 0) return;
        active = true;
        bind_this_thread(cpus);
    }
    ~ThreadBind() { if (active) (void)pthread_setaffinity_np(pthread_self(), sizeof saved, &saved); }
};

// temporary device buffer of the synchronous tap / stage entry points: freed on every return path
template <class Tp>
struct DevBuf {
    Tp *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count) { return cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(Tp)); }
    operator Tp *() const { return p; }
};

// several device threads of one call may fail at once (for_each_slot_range): the message is written under the handle's mutex
void set_error(npswf_handle *h, const char *msg)
{
    std::lock_guard<std::mutex> lk(h->mu);
    h->err = msg;
}

#define CU_TRY(h, expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            char _b[512];                                                                            \
            snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            set_error(h, _b);                                                                        \
            return NPSWF_ERR_CUDA;                                                                   \
        }                                                                                            \
    } while (0)

// Natural cubic spline through the block's knots as GSL builds it for ROOT::Math::Interpolator(kCSPLINE) (T2:612-619;
// gsl_interp_cspline: cspline_init + gsl_linalg_solve_symm_tridiag): c_0 = c_{n-1} = 0, interior c_i from
// the symmetric tridiagonal system factorised as L D L^T in GSL's operation order, so the coefficients carry the
// bits the reference's spline has.  coef[i] = {y_i, b_i, c_i, d_i} for the interval [x_i, x_{i+1}].
void build_spline(const double *x, const double *y, double *coef)
{
    const int n = T, sys = n - 2;
    std::vector<double> c(n, 0.0), off(sys), diag(sys), rhs(sys), alpha(sys), gamma(sys), z(sys), sol(sys);
    for (int i = 0; i < sys; i++) {
        const double h0 = x[i + 1] - x[i], h1 = x[i + 2] - x[i + 1];
        const double dy0 = y[i + 1] - y[i], dy1 = y[i + 2] - y[i + 1];
        const double r0 = (h0 != 0.0) ? 1.0 / h0 : 0.0, r1 = (h1 != 0.0) ? 1.0 / h1 : 0.0;
        off[i] = h1;
        diag[i] = 2.0 * (h1 + h0);
        rhs[i] = 3.0 * (dy1 * r1 - dy0 * r0);
    }
    alpha[0] = diag[0];
    gamma[0] = off[0] / alpha[0];
    for (int i = 1; i < sys - 1; i++) {
        alpha[i] = diag[i] - off[i - 1] * gamma[i - 1];
        gamma[i] = off[i] / alpha[i];
    }
    if (sys > 1) alpha[sys - 1] = diag[sys - 1] - off[sys - 2] * gamma[sys - 2];
    z[0] = rhs[0];
    for (int i = 1; i < sys; i++) z[i] = rhs[i] - gamma[i - 1] * z[i - 1];
    for (int i = 0; i < sys; i++) sol[i] = z[i] / alpha[i];
    c[sys] = sol[sys - 1];
    for (int i = sys - 2; i >= 0; i--) c[1 + i] = sol[i] - gamma[i] * c[2 + i];
    for (int i = 0; i < n - 1; i++) {
        const double h = x[i + 1] - x[i], dy = y[i + 1] - y[i];
        coef[4 * i + 0] = y[i];
        coef[4 * i + 1] = (dy / h) - h * (c[i + 1] + 2.0 * c[i]) / 3.0;
        coef[4 * i + 2] = c[i];
        coef[4 * i + 3] = (c[i + 1] - c[i]) / (3.0 * h);
    }
}

template <class Tp>
int dev_alloc(npswf_handle *h, DevSlot &s, Tp **p, size_t count)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(Tp));
    if (e != cudaSuccess) {
        set_error(h, (std::string("cudaMalloc failed: ") + cudaGetErrorString(e)).c_str());
        return NPSWF_ERR_NOMEM;
    }
    s.owned.push_back(q);
    *p = reinterpret_cast<Tp *>(q);
    return 0;
}

template <class Tp>
int dev_upload(npswf_handle *h, DevSlot &s, const Tp **dst, const Tp *src, size_t count)
{
    Tp *p = nullptr;
    int rc = dev_alloc(h, s, &p, count);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpy(p, src, count * sizeof(Tp), cudaMemcpyHostToDevice));
    *dst = p;
    return 0;
}

template <class Tp>
void dev_free(DevSlot &s, Tp **p)
{
    if (!*p) return;
    for (size_t i = 0; i < s.owned.size(); i++)
        if (s.owned[i] == (void *)*p) { s.owned.erase(s.owned.begin() + (long)i); break; }
    cudaFree((void *)*p);
    *p = nullptr;
}

// scratch buffers and job lists of a workspace for `cap` events
int alloc_scratch(npswf_handle *h, DevSlot &s, Workspace &w, int64_t cap)
{
    w.cap = cap;
    const size_t nb = (size_t)cap * B;
    int rc = 0;
    if ((rc = dev_alloc(h, s, &w.mf, nb * T))) return rc;
    if ((rc = dev_alloc(h, s, &w.minsig, nb))) return rc;
    if ((rc = dev_alloc(h, s, &w.flags, nb))) return rc;
    if ((rc = dev_alloc(h, s, &w.cont_list, (size_t)6 * nb))) return rc;
    if ((rc = dev_alloc(h, s, &w.cont_state, (size_t)3 * nb * FT_CONT_STRIDE))) return rc;
    if ((rc = dev_alloc(h, s, &w.fit_list, (size_t)(MAXP + 1) * nb))) return rc;
    if ((rc = dev_alloc(h, s, &w.fit_dense, (size_t)(MAXP + 1) * nb))) return rc;
    return 0;
}

// The device path cuts a call into two chunks of up to dev_cap events: the scratch of both workspaces grows to the
// chunk size a call needs the first time it is needed (device idle), so a handle that only sees small calls stays small.
int grow_scratch(npswf_handle *h, DevSlot &s, int64_t cap)
{
    if (s.ws[0].cap >= cap && s.ws[1].cap >= cap) return 0;
    CU_TRY(h, cudaDeviceSynchronize());
    for (int i = 0; i < 2; i++) {
        Workspace &w = s.ws[i];
        if (w.cap >= cap) continue;
        dev_free(s, &w.mf); dev_free(s, &w.minsig); dev_free(s, &w.flags); dev_free(s, &w.cont_list);
        dev_free(s, &w.cont_state); dev_free(s, &w.fit_list); dev_free(s, &w.fit_dense);
        int rc = alloc_scratch(h, s, w, cap);
        if (rc) return rc;
    }
    return 0;
}

int alloc_workspace(npswf_handle *h, DevSlot &s, Workspace &w, int64_t cap, bool io)
{
    w.cap = cap;
    w.io = io;
    const size_t nb = (size_t)cap * B;
    int rc = 0;
    if (io) {
        if ((rc = dev_alloc(h, s, &w.signal, nb * T))) return rc;
        if ((rc = dev_alloc(h, s, &w.pres, nb))) return rc;
        if ((rc = dev_alloc(h, s, &w.corr, (size_t)cap))) return rc;
        if ((rc = dev_alloc(h, s, &w.wfnpulse, nb))) return rc;
        if ((rc = dev_alloc(h, s, &w.wftime, nb * MAXP))) return rc;
        if ((rc = dev_alloc(h, s, &w.wfampl, nb * MAXP))) return rc;
        if ((rc = dev_alloc(h, s, &w.chi2, nb))) return rc;
        if ((rc = dev_alloc(h, s, &w.timewf, nb))) return rc;
        if ((rc = dev_alloc(h, s, &w.amplwf, nb))) return rc;
        if ((rc = dev_alloc(h, s, &w.status, nb))) return rc;
        if ((rc = dev_alloc(h, s, &w.mask, nb))) return rc;
    }
    if ((rc = alloc_scratch(h, s, w, cap))) return rc;
    if ((rc = dev_alloc(h, s, &w.fit_count, 128))) return rc;
    if ((rc = dev_alloc(h, s, &w.bucket_count, (size_t)(MAXP + 1) * B))) return rc;
    CU_TRY(h, cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    CU_TRY(h, cudaEventCreateWithFlags(&w.ev_in, cudaEventDisableTiming));
    CU_TRY(h, cudaEventCreateWithFlags(&w.ev_cmp, cudaEventDisableTiming));
    CU_TRY(h, cudaEventCreateWithFlags(&w.ev_out, cudaEventDisableTiming));
    return 0;
}

// Lazily allocate the staging buffers used by the host-buffer entry points.
int ensure_io(npswf_handle *h, DevSlot &s)
{
    for (int i = 0; i < 3; i++) {
        Workspace &w = s.ws[i];
        if (w.io) continue;
        if (!w.stream) {   // the third workspace exists only for the host pipeline
            int rc0 = alloc_workspace(h, s, w, h->chunk, false);
            if (rc0) return rc0;
        }
        w.io_cap = std::min<int64_t>(w.cap, h->chunk);
        const size_t nb = (size_t)w.io_cap * B;
        int rc = 0;
        if ((rc = dev_alloc(h, s, &w.signal, nb * T))) return rc;
        if ((rc = dev_alloc(h, s, &w.pres, nb))) return rc;
        if ((rc = dev_alloc(h, s, &w.corr, (size_t)w.io_cap))) return rc;
        if ((rc = dev_alloc(h, s, &w.wfnpulse, nb))) return rc;
        if ((rc = dev_alloc(h, s, &w.wftime, nb * MAXP))) return rc;
        if ((rc = dev_alloc(h, s, &w.wfampl, nb * MAXP))) return rc;
        if ((rc = dev_alloc(h, s, &w.chi2, nb))) return rc;
        if ((rc = dev_alloc(h, s, &w.timewf, nb))) return rc;
        if ((rc = dev_alloc(h, s, &w.amplwf, nb))) return rc;
        if ((rc = dev_alloc(h, s, &w.status, nb))) return rc;
        if ((rc = dev_alloc(h, s, &w.mask, nb))) return rc;
        w.io = true;
    }
    return 0;
}

__global__ void widen_counts_kernel(const int16_t *__restrict__ c, double *__restrict__ out, double lsb, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __dmul_rn((double)c[i], lsb);
}

// job lists from an explicit mask (npswf_fitwf_batch): one thread per (event, block)
__global__ void build_jobs_kernel(const uint8_t *__restrict__ mask, const int32_t *__restrict__ wfnpulse, long long n_items,
                                  int *__restrict__ bucket_count, int *__restrict__ bucket_list, int bucket_cap,
                                  double *__restrict__ chi2, uint8_t *__restrict__ status)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    chi2[i] = -100.0;
    status[i] = 0;
    int n = wfnpulse[i];
    if (n > MAXP) n = MAXP;
    if (mask[i] && n > 0) {
        const int bucket = n * B + (int)(i % B);
        const int idx = atomicAdd(&bucket_count[bucket], 1);
        bucket_list[(size_t)bucket * bucket_cap + idx] = (int)i;
    }
}

// Buckets -> dense, exactly block-major job lists (one per multiplicity) + their lengths.  CTA (c, n) owns the
// blocks [27c, 27c + 27) of multiplicity n + 1: offset = sum of the counts of the blocks before them.
constexpr int FC_BLOCKS = 27;
__global__ void __launch_bounds__(256)
fit_compact_kernel(const int *__restrict__ bucket_count, const int *__restrict__ bucket_list, int bucket_cap,
                   int *__restrict__ fit_count, int *__restrict__ fit_dense, long long dense_stride)
{
    const int n = blockIdx.y + 1;
    const int b0 = blockIdx.x * FC_BLOCKS;
    __shared__ int s_part[8];
    int part = 0;
    for (int b = threadIdx.x; b < b0; b += blockDim.x) part += bucket_count[n * B + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    int off = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) off += s_part[i];
    const int base = off;
    int *dst = fit_dense + (size_t)n * dense_stride;
    for (int b = b0; b < b0 + FC_BLOCKS && b < B; b++) {
        const int c = bucket_count[n * B + b];
        const int *src = bucket_list + (size_t)(n * B + b) * bucket_cap;
        for (int i = threadIdx.x; i < c; i += blockDim.x) dst[off + i] = src[i];
        off += c;
    }
    if (threadIdx.x == 0 && off > base) atomicAdd(&fit_count[n], off - base);
}

int launch_compact(npswf_handle *h, cudaStream_t st, Workspace &w)
{
    fit_compact_kernel<<<dim3((B + FC_BLOCKS - 1) / FC_BLOCKS, MAXP), 256, 0, st>>>(w.bucket_count, w.fit_list, (int)w.cap,
                                                                                     w.fit_count, w.fit_dense,
                                                                                     (long long)w.cap * B);
    CU_TRY(h, cudaGetLastError());
    return 0;
}

// search kernel launch: persistent CTAs, 32 spectra per CTA batch
int launch_search(npswf_handle *h, DevSlot &s, cudaStream_t st, SearchArgs &a)
{
    a.kp = h->kp;
    a.gold1 = s.gold1;
    const long long batches = (a.n_items + SRB - 1) / SRB;
    const int grid = (int)std::min<long long>(batches, (long long)s.sm_count * s.occ_search);
    if (grid <= 0) return 0;
    search_kernel<<<grid, SEARCH_THREADS, SEARCH_SMEM, st>>>(a);
    CU_TRY(h, cudaGetLastError());
    return 0;
}

int launch_front(npswf_handle *h, DevSlot &s, cudaStream_t st, const double *sig, const int32_t *pres, int64_t n,
                 float *mf, double *minsig, uint8_t *flags, int do_mf, int do_thr)
{
    const int grid = (int)std::min<int64_t>(n, (int64_t)s.sm_count * s.occ_front);
    front_kernel<<<grid, FRONT_THREADS, FRONT_SMEM, st>>>(sig, pres, n, s.cal, h->kp, mf, minsig, flags, do_mf, do_thr);
    CU_TRY(h, cudaGetLastError());
    return 0;
}

int launch_fits(npswf_handle *h, DevSlot &s, cudaStream_t st, Workspace &w, const double *sig, const double *corr,
                double *wftime, double *wfampl, double *chi2, double *timewf, double *amplwf, uint8_t *status)
{
    const long long stride = (long long)w.cap * B;
    // The per-multiplicity kernels are independent (disjoint jobs) and every one of them ends in a tail of a
    // few long fits, so they run on four side streams forked from / joined to the caller's stream.
    cudaStream_t caller = st;
    const bool conc = s.fit_concurrent && !h->profiling;
    if (conc) {
        CU_TRY(h, cudaEventRecord(s.fit_fork, caller));
        for (int i = 0; i < 4; i++) CU_TRY(h, cudaStreamWaitEvent(s.fit_stream[i], s.fit_fork, 0));
    }
    for (int N = 1; N <= MAXP; N++) {
        // N = 1, 2, 3 have a stream each; N >= 4 (few jobs, long tails) are dealt round the four streams so that their
        // tails overlap instead of queueing behind one another
        if (conc) st = s.fit_stream[(N - 1) % 4];
        const int *list = w.fit_dense + (size_t)N * stride;
        const int *cnt = w.fit_count + N;
        int *next = w.fit_count + 16 + N;  // per-multiplicity job cursor, zeroed with fit_count
        const bool vm_fast = h->fit_mode == NPSWF_FIT_VM && N <= NPSWF_VM_MAXN && h->unit_knots && s.occ_vm_thread[N] > 0;
        if (h->fit_mode == NPSWF_FIT_MIGRAD || (h->fit_mode == NPSWF_FIT_VM && !vm_fast)) {
            // the reference's own minimiser (Migrad, numerical gradients, strategy 1 -> 2), one warp per fit
            // (VM mode: 7+ pulses and general knots have no analytic-path kernel and take the exact one)
            MigradArgs ma{list, cnt, next, N, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr};
            const int cls = migrad_class(N);
            if (N <= 3 && s.migrad_thread && s.occ_migrad_thread[N] > 0 && h->unit_knots) {
                // one thread per fit for traces on the ADC lattice; anything else is handed to the warp-per-fit kernel
                if (s.mg_wtab_lsb != h->pack_lsb) {   // (re)built on the stream that uses it first; later users are ordered behind it by the fork
                    CU_TRY(h, migrad_build_wtab(s.mg_wtab, h->pack_lsb, caller));
                    s.mg_wtab_lsb = h->pack_lsb;
                    s.mg_wlow = migrad_wtab_floor(h->pack_lsb);
                    if (conc) {
                        CU_TRY(h, cudaEventRecord(s.fit_fork, caller));
                        for (int i = 0; i < 4; i++) CU_TRY(h, cudaStreamWaitEvent(s.fit_stream[i], s.fit_fork, 0));
                    }
                }
                ma.wtab = s.mg_wtab; ma.lsb = h->pack_lsb; ma.wlow = s.mg_wlow;
                ma.ho_count = w.fit_count + 32 + N;
                ma.ho_list = w.cont_list + (size_t)(N - 1) * stride;
                CU_TRY(h, migrad_thread_launch(N, s.sm_count * s.occ_migrad_thread[N], st, ma));
                MigradArgs mb = ma;
                mb.job_list = ma.ho_list; mb.job_count = ma.ho_count; mb.job_next = w.fit_count + 48 + N;
                CU_TRY(h, migrad_launch(cls, s.sm_count * s.occ_migrad[cls], st, mb));
            } else {
                CU_TRY(h, migrad_launch(cls, s.sm_count * s.occ_migrad[cls], st, ma));
            }
        } else if (vm_fast) {
            // Migrad's own recursion with analytic derivatives (kernel_fit_vm.cuh); what leaves the common path goes to
            // the exact warp-per-fit Migrad kernel
            int *ccnt = w.fit_count + 32 + N, *cnext = w.fit_count + 48 + N;
            int *clist = w.cont_list + (size_t)(N - 1) * stride;
            const int vgrid = s.sm_count * s.occ_vm_thread[N];
#define NPSWF_VM_LAUNCH(NN)                                                                                                    \
    fit_vm_thread_kernel<NN><<<vgrid, VM_THREADS, VM_SMEM, st>>>(list, cnt, next, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, \
                                                                 timewf, amplwf, status, s.ctr, ccnt, clist)
            switch (N) {
            case 1: NPSWF_VM_LAUNCH(1); break;
            case 2: NPSWF_VM_LAUNCH(2); break;
            case 3: NPSWF_VM_LAUNCH(3); break;
#if NPSWF_VM_MAXN >= 6
            case 4: NPSWF_VM_LAUNCH(4); break;
            case 5: NPSWF_VM_LAUNCH(5); break;
            default: NPSWF_VM_LAUNCH(6); break;
#else
            default: break;
#endif
            }
#undef NPSWF_VM_LAUNCH
            CU_TRY(h, cudaGetLastError());
            // the hand-over lists are short (~1 % of the fits): the warp-per-fit kernel finishes a fit in a fraction of a
            // millisecond, a thread of the thread-per-fit kernel needs 3-50 ms for one, which would be the tail of the stage
            MigradArgs ma{clist, ccnt, cnext, N, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr};
            const int cls = migrad_class(N);
            CU_TRY(h, migrad_launch(cls, s.sm_count * s.occ_migrad[cls], st, ma));
        } else if (!h->unit_knots) {
            // general interpX: the warp-per-fit LM kernel with a bisection per spline evaluation (both attempts)
            if (N <= 6)
                fit_kernel<13><<<s.sm_count * s.occ_fit_mid, FIT_THREADS, sizeof(FitSmem<13>) * FIT_WARPS, st>>>(
                    list, cnt, N, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr, 0, 0, next);
            else
                fit_kernel<25><<<s.sm_count * s.occ_fit_big, FIT_THREADS, sizeof(FitSmem<25>) * FIT_WARPS, st>>>(
                    list, cnt, N, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr, 0, 0, next);
        } else if (N <= 3 && s.fit_thread) {
            // thread-per-fit for the first tries of every fit, then the sub-warp kernel on the (rare) fits handed over
            int *ccnt = w.fit_count + 32 + N, *cnext = w.fit_count + 48 + N;
            int *clist = w.cont_list + (size_t)(N - 1) * stride;
            double *cstate = w.cont_state + (size_t)(N - 1) * stride * FT_CONT_STRIDE;
            const int tgrid = s.sm_count * s.occ_fit_thread[N], sgrid = s.sm_count * 3;
            if (N == 1) {
                fit_thread_kernel<1><<<tgrid, FT_THREADS, ft_smem(1), st>>>(
                    list, cnt, next, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr, ccnt, clist, cstate);
                CU_TRY(h, cudaGetLastError());
                fit_small_kernel<1, 16, 3><<<sgrid, FS_THREADS, 0, st>>>(
                    clist, ccnt, cnext, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr, cstate);
            } else if (N == 2) {
                fit_thread_kernel<2><<<tgrid, FT_THREADS, ft_smem(2), st>>>(
                    list, cnt, next, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr, ccnt, clist, cstate);
                CU_TRY(h, cudaGetLastError());
                fit_small_kernel<2, 16, 3><<<sgrid, FS_THREADS, 0, st>>>(
                    clist, ccnt, cnext, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr, cstate);
            } else {
                fit_thread_kernel<3><<<tgrid, FT_THREADS, ft_smem(3), st>>>(
                    list, cnt, next, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr, ccnt, clist, cstate);
                CU_TRY(h, cudaGetLastError());
                fit_small_kernel<3, 16, FS_MINB3><<<s.sm_count * s.occ_fit_small[3], FS_THREADS, 0, st>>>(
                    clist, ccnt, cnext, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr, cstate);
            }
        } else if (N >= 4 && N <= 6 && s.fit_thread) {
            // thread-per-fit also for 4-6 pulses (the normal equations spill to L1-resident local memory, which is
            // read once per try); it runs the whole first attempt, the warp-per-fit kernel the retries
            int *ccnt = w.fit_count + 32 + N;
            int *clist = w.cont_list + (size_t)(N - 1) * stride;
            const int tgrid = s.sm_count * s.occ_fit_thread[4];
            if (N == 4)
                fit_thread_kernel<4><<<tgrid, FT_THREADS, ft_smem(4), st>>>(list, cnt, next, sig, corr, s.cal, h->kp, wftime, wfampl,
                                                                        chi2, timewf, amplwf, status, s.ctr, ccnt, clist, nullptr);
            else if (N == 5)
                fit_thread_kernel<5><<<tgrid, FT_THREADS, ft_smem(5), st>>>(list, cnt, next, sig, corr, s.cal, h->kp, wftime, wfampl,
                                                                        chi2, timewf, amplwf, status, s.ctr, ccnt, clist, nullptr);
            else
                fit_thread_kernel<6><<<tgrid, FT_THREADS, ft_smem(6), st>>>(list, cnt, next, sig, corr, s.cal, h->kp, wftime, wfampl,
                                                                        chi2, timewf, amplwf, status, s.ctr, ccnt, clist, nullptr);
            CU_TRY(h, cudaGetLastError());
            // P <= 13: half the shared memory of the 25-parameter instance, twice the resident warps; jobs claimed one by one
            fit_kernel<13><<<s.sm_count * s.occ_fit_mid, FIT_THREADS, sizeof(FitSmem<13>) * FIT_WARPS, st>>>(
                clist, ccnt, N, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr, 1, h->kp.fit_max_iter,
                w.fit_count + 48 + N);
        } else if (N == 1) {
            fit_small_kernel<1, 8, FS_MINB1><<<s.sm_count * s.occ_fit_small[1], FS_THREADS, 0, st>>>(
                list, cnt, next, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr);
        } else if (N == 2) {
            fit_small_kernel<2, 8, FS_MINB2><<<s.sm_count * s.occ_fit_small[2], FS_THREADS, 0, st>>>(
                list, cnt, next, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr);
        } else if (N == 3) {
            fit_small_kernel<3, 16, FS_MINB3><<<s.sm_count * s.occ_fit_small[3], FS_THREADS, 0, st>>>(
                list, cnt, next, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr);
        } else {
            fit_kernel<25><<<s.sm_count * s.occ_fit_big, FIT_THREADS, sizeof(FitSmem<25>) * FIT_WARPS, st>>>(
                list, cnt, N, sig, corr, s.cal, h->kp, wftime, wfampl, chi2, timewf, amplwf, status, s.ctr, 0, 0, next);
        }
        CU_TRY(h, cudaGetLastError());
    }
    if (conc) {
        for (int i = 0; i < 4; i++) {
            CU_TRY(h, cudaEventRecord(s.fit_join[i], s.fit_stream[i]));
            CU_TRY(h, cudaStreamWaitEvent(caller, s.fit_join[i], 0));
        }
    }
    return 0;
}

// The whole per-chunk pipeline on device pointers: front -> search -> fits.
int run_chunk(npswf_handle *h, DevSlot &s, Workspace &w, cudaStream_t st, int64_t n, const double *sig,
              const int32_t *pres, const double *corr, int32_t *wfnpulse, double *wftime, double *wfampl, double *chi2,
              double *timewf, double *amplwf, uint8_t *status)
{
    CU_TRY(h, cudaMemsetAsync(w.fit_count, 0, 128 * sizeof(int), st));
    CU_TRY(h, cudaMemsetAsync(w.bucket_count, 0, (size_t)(MAXP + 1) * B * sizeof(int), st));
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (h->profiling) {
        for (int i = 0; i < 4; i++) {
            if (!s.prof_pool.empty()) { ev[i] = s.prof_pool.back(); s.prof_pool.pop_back(); }
            else CU_TRY(h, cudaEventCreate(&ev[i]));
        }
        CU_TRY(h, cudaEventRecord(ev[0], st));
    }
    int rc = launch_front(h, s, st, sig, pres, n, w.mf, w.minsig, w.flags, 1, 1);
    if (rc) return rc;
    if (h->profiling) CU_TRY(h, cudaEventRecord(ev[1], st));
    SearchArgs sa{};
    sa.hist = w.mf; sa.flags = w.flags; sa.minsig = w.minsig; sa.signal = sig; sa.n_items = (long long)n * B;
    sa.wfnpulse = wfnpulse; sa.wftime = wftime; sa.wfampl = wfampl; sa.chi2 = chi2; sa.timewf = timewf; sa.amplwf = amplwf;
    sa.status = status; sa.bucket_count = w.bucket_count; sa.bucket_list = w.fit_list; sa.bucket_cap = (int)w.cap;
    sa.ctr = s.ctr;
    if ((rc = launch_search(h, s, st, sa))) return rc;
    if ((rc = launch_compact(h, st, w))) return rc;
    if (h->profiling) CU_TRY(h, cudaEventRecord(ev[2], st));
    rc = launch_fits(h, s, st, w, sig, corr, wftime, wfampl, chi2, timewf, amplwf, status);
    if (rc) return rc;
    if (h->profiling) {
        CU_TRY(h, cudaEventRecord(ev[3], st));
        for (int i = 0; i < 4; i++) s.prof_events.push_back(ev[i]);
    }
    return 0;
}

// Resolve the recorded stage events of a slot (the stream must have been synchronised).
int fold_profile(npswf_handle *h, DevSlot &s)
{
    for (size_t i = 0; i + 3 < s.prof_events.size(); i += 4) {
        float ms[3] = {0, 0, 0};
        for (int k = 0; k < 3; k++) CU_TRY(h, cudaEventElapsedTime(&ms[k], s.prof_events[i + k], s.prof_events[i + k + 1]));
        {
            std::lock_guard<std::mutex> lk(h->mu);   // one thread per device folds into the handle's totals
            for (int k = 0; k < 3; k++) h->stage_ms[k] += ms[k];
            h->stage_chunks++;
        }
        for (int k = 0; k < 4; k++) s.prof_pool.push_back(s.prof_events[i + k]);
    }
    s.prof_events.clear();
    return 0;
}

template <class F>
int for_each_slot_range(npswf_handle *h, int64_t n_events, F &&fn)
{
    const int nd = (int)h->slots.size();
    if (nd == 1) return fn(0, (int64_t)0, n_events);
    std::vector<int> rcs(nd, 0);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++) {
        const int64_t lo = n_events * d / nd, hi = n_events * (d + 1) / nd;  // contiguous event ranges
        th.emplace_back([&, d, lo, hi]() { rcs[d] = (hi > lo) ? fn(d, lo, hi) : 0; });
    }
    for (auto &t : th) t.join();
    for (int rc : rcs)
        if (rc) return rc;
    return 0;
}

struct HostIO {
    const double *signal = nullptr;
    const int16_t *counts = nullptr;
    double lsb = 0;
    const double *packed = nullptr;       // packed stream + event offsets instead of signal / pres
    const int64_t *offsets = nullptr;
    const int32_t *pres = nullptr;
    const double *corr = nullptr;
    int32_t *wfnpulse = nullptr;
    double *wftime = nullptr, *wfampl = nullptr, *chi2 = nullptr, *timewf = nullptr, *amplwf = nullptr;
    uint8_t *status = nullptr;
    // flat outputs (npswf_analyze_batch_flat)
    bool flat = false;
    int64_t *pulse_offset = nullptr;
    int32_t *pulse_count = nullptr;
    double *pool_t = nullptr, *pool_a = nullptr;
    int64_t pool_cap = 0, n_total = 0;
    int64_t *pulses_out = nullptr;   // [n_devices]: pulses written by each device range
};

// Every call works on per-slot state shared by all calls (scratch, job lists and cursors, fit streams, fork / join
// events).  The device-path call is asynchronous on the caller's stream, so it leaves an event behind; a host-buffer
// call waits for it on the host, the next device-path call makes its stream wait for it -- calls on different
// streams are thereby serialised instead of corrupting one another's job lists.
int wait_last_use(npswf_handle *h, DevSlot &s)
{
    if (s.last_use_valid) CU_TRY(h, cudaEventSynchronize(s.last_use));
    return 0;
}

int analyze_range_impl(npswf_handle *h, int d, int64_t lo, int64_t hi, const HostIO &io);

// Chunked, double-buffered host pipeline on one device for events [lo, hi).  On any error nothing is left in flight
// that still reads or writes the caller's buffers.
int analyze_range(npswf_handle *h, int d, int64_t lo, int64_t hi, const HostIO &io)
{
    const int rc = analyze_range_impl(h, d, lo, hi, io);
    if (rc) {
        cudaSetDevice(h->slots[d].device);
        cudaDeviceSynchronize();
        (void)cudaGetLastError();
    }
    return rc;
}

int analyze_range_impl(npswf_handle *h, int d, int64_t lo, int64_t hi, const HostIO &io)
{
    DevSlot &s = h->slots[d];
    CU_TRY(h, cudaSetDevice(s.device));
    int rc = wait_last_use(h, s);
    if (rc) return rc;
    if ((rc = ensure_io(h, s))) return rc;
    ThreadBind bind(s.local_cpus);   // this thread packs a share of every chunk and first-touches the staging buffers
    // Host buffers: a three-stage pipeline over the two workspaces -- uploads on their own stream, every kernel on
    // the chunk's own stream, downloads on a
    // third stream -- so the copy engines of both directions run under the kernels of the neighbouring chunks.
    // Eight to twelve chunks per call (multiples of 148 events, at least 296) keep the uncovered first upload and
    // last download short.
    // (the binary64 layouts are bound by the upload: more, smaller chunks; the int16 layout by the kernels: fewer, larger)
    const int64_t target = io.counts ? 8 : 12;
    int64_t chunk = h->chunk;
    if (hi - lo > 2 * 296)
        chunk = std::min<int64_t>(h->chunk, std::max<int64_t>(296, ((hi - lo + target * 148 - 1) / (target * 148)) * 148));
    cudaStream_t s_in = s.copy_in, s_out = s.copy_out;
    // binary64 host layout: events go over as int16 counts when that is lossless (host_pack.hpp).  The host threads
    // (pack_rate, measured) and the copy engine (~48 GB/s from pinned memory) feed the device side by side: in auto
    // mode the first n_raw events of every chunk are uploaded as they are while the host packs the rest, with the
    // split chosen so that both finish together -- (1-f) / pack_rate = (f + (1-f)/4) / dma_rate.  From pageable memory
    // the raw upload is slow and blocks the caller, so everything is packed.
    bool pack = io.signal && !io.counts && !io.packed && h->pack_mode != 0 && !s.pack_unavailable;
    bool pinned_src = false;
    if (pack) {
        if (!s.packer) s.packer.reset(new PackPool(h->pack_threads, s.local_cpus));
        const size_t need = (size_t)std::min<int64_t>(chunk, hi - lo) * B * T;
        if (need > s.stage_cap) {
            s.stage_cap = 0;
            for (int i = 0; i < 3 && pack; i++) {
                if (s.stage_busy[i]) CU_TRY(h, cudaEventSynchronize(s.stage_ev[i]));
                s.stage_busy[i] = false;
                if (s.stage[i]) CU_TRY(h, cudaFreeHost(s.stage[i]));
                s.stage[i] = nullptr;
                if (!s.stage_ev[i]) CU_TRY(h, cudaEventCreateWithFlags(&s.stage_ev[i], cudaEventDisableTiming));
                if (cudaHostAlloc((void **)&s.stage[i], need * sizeof(int16_t), cudaHostAllocPortable) != cudaSuccess) {
                    // no pinned memory for the staging buffers: this is a transport optimisation, not a requirement --
                    // the traces go over as the caller's doubles from now on
                    (void)cudaGetLastError();
                    s.stage[i] = nullptr;
                    for (int j = 0; j < i; j++) { cudaFreeHost(s.stage[j]); s.stage[j] = nullptr; }
                    s.pack_unavailable = true;
                    pack = false;
                }
            }
            if (pack) s.stage_cap = need;
        }
    }
    if (pack) {
        cudaPointerAttributes at{};
        pinned_src = cudaPointerGetAttributes(&at, io.signal) == cudaSuccess && at.type == cudaMemoryTypeHost;
        (void)cudaGetLastError();
        if (s.pack_rate <= 0) s.pack_rate = 4.0e9 * s.packer->threads();
    }
    // flat outputs: this range owns the slice [pool_lo, pool_hi) of the caller's pulse pools; the copy of a chunk's
    // pulses is enqueued one iteration later, when its pulse count has reached the host (no bubble in the pipeline)
    int64_t pool_cur = 0, pool_hi = 0;
    struct Pending { Workspace *w; int64_t e0, n; };
    Pending pend{nullptr, 0, 0}, pend2{nullptr, 0, 0};   // the two chunks whose pulse copies are not enqueued yet (older first)
    if (io.flat) {
        pool_cur = (int64_t)((__int128)io.pool_cap * lo / io.n_total);
        pool_hi = (int64_t)((__int128)io.pool_cap * hi / io.n_total);
        for (int i = 0; i < 3; i++) {
            Workspace &w = s.ws[i];
            if (w.flat_t) continue;
            if ((rc = dev_alloc(h, s, &w.ev_total, (size_t)w.cap))) return rc;
            if ((rc = dev_alloc(h, s, &w.ev_off, (size_t)w.cap + 1))) return rc;
            if ((rc = dev_alloc(h, s, &w.flat_t, (size_t)w.io_cap * B * MAXP))) return rc;
            if ((rc = dev_alloc(h, s, &w.flat_a, (size_t)w.io_cap * B * MAXP))) return rc;
            CU_TRY(h, cudaHostAlloc((void **)&w.h_ev_off, ((size_t)w.cap + 1) * sizeof(int), cudaHostAllocPortable));
            CU_TRY(h, cudaEventCreateWithFlags(&w.ev_tot, cudaEventDisableTiming));
        }
    }
    auto finish_flat = [&]() -> int {
        if (!pend.w) return 0;
        Workspace &pw = *pend.w;
        CU_TRY(h, cudaEventSynchronize(pw.ev_tot));
        const int64_t total = pw.h_ev_off[pend.n];
        if (pool_cur + total > pool_hi) {
            char b[256];
            snprintf(b, sizeof b, "npswf_analyze_batch_flat: pulse pool too small (events %lld..%lld need %lld more pulses, %lld left in "
                     "this range's share)", (long long)pend.e0, (long long)(pend.e0 + pend.n), (long long)total, (long long)(pool_hi - pool_cur));
            set_error(h, b);
            cudaDeviceSynchronize();   // leave nothing in flight that still reads or writes the caller's buffers
            pend.w = nullptr;
            pend2.w = nullptr;
            return NPSWF_ERR_NOMEM;
        }
        if (total > 0) {
            CU_TRY(h, cudaMemcpyAsync(io.pool_t + pool_cur, pw.flat_t, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, s_out));
            CU_TRY(h, cudaMemcpyAsync(io.pool_a + pool_cur, pw.flat_a, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, s_out));
        }
        for (int64_t i = 0; i < pend.n; i++) {
            if (io.pulse_offset) io.pulse_offset[pend.e0 + i] = pool_cur + pw.h_ev_off[i];
            if (io.pulse_count) io.pulse_count[pend.e0 + i] = pw.h_ev_off[i + 1] - pw.h_ev_off[i];
        }
        pool_cur += total;
        CU_TRY(h, cudaEventRecord(pw.ev_out, s_out));
        pend = pend2;
        pend2.w = nullptr;
        return 0;
    };
    int which = 0;
    int64_t k = 0;
    // The pipeline is empty while the first chunk is packed and uploaded and while the last one is downloaded: the
    // first chunks ramp up (a quarter, then half of the regular size) and the last one is half a chunk when the call
    // is long enough for that to pay (a small chunk uses the kernels less well).
    std::vector<int64_t> cuts;   // chunk k = events [cuts[k], cuts[k+1])
    {
        const int64_t unit = 148, q = std::max<int64_t>(unit, chunk / 4 / unit * unit), hf = std::max<int64_t>(unit, chunk / 2 / unit * unit);
        int64_t e = lo;
        cuts.push_back(e);
        if (hi - lo >= 6 * chunk && chunk >= 2 * unit && h->chunk_ramp) {
            e += q; cuts.push_back(e);
            e += hf; cuts.push_back(e);
            while (hi - e > chunk + hf) { e += chunk; cuts.push_back(e); }
            if (hi - e > hf) { e = hi - hf; cuts.push_back(e); }
        } else {
            while (hi - e > chunk) { e += chunk; cuts.push_back(e); }
        }
        cuts.push_back(hi);
    }
    for (size_t ci = 0; ci + 1 < cuts.size(); ci++, which = (which + 1) % 3, k++) {
        const int64_t e0 = cuts[ci];
        Workspace &w = s.ws[which];
        // Three workspaces, each with its own stream, two chunks computing at a time: the front and search kernels of
        // chunk k+1 fill the SMs that the fit tails of chunk k leave idle (as in the device path), while the third
        // workspace is being downloaded / refilled -- with two workspaces the later completion of every chunk stalls
        // the uploads (measured: 61.8 instead of 54.2 ms per 4 736 events).  Stage profiling: one stream, clean times.
        cudaStream_t s_cmp = h->profiling ? s.ws[0].stream : w.stream;
        const int64_t n = cuts[ci + 1] - e0;
        const size_t nb = (size_t)n * B, ob = (size_t)e0 * B;
        // events [0, n_raw) of the chunk travel as doubles, [n_raw, n) as counts
        int64_t n_raw = n;
        if (pack) {
            double f = 0.0;
            if (h->pack_mode == 1 && pinned_src) {
                if (s.dma_pending && cudaEventQuery(s.dma_ev[1]) == cudaSuccess) {   // fold the last timed raw upload in
                    float ms = 0;
                    if (cudaEventElapsedTime(&ms, s.dma_ev[0], s.dma_ev[1]) == cudaSuccess && ms > 0)
                        s.dma_rate = 0.5 * s.dma_rate + 0.5 * s.dma_bytes / (ms * 1e-3);
                    s.dma_pending = false;
                }
                (void)cudaGetLastError();
                const double a = s.dma_rate / s.pack_rate;
                f = std::min(1.0, std::max(0.0, (a - 0.25) / (a + 0.75)));
                if (f > 0.9) f = 1.0;
                // a packer far slower than the copy engine is short of cores or of memory bandwidth (several ranks on
                // one host): its reads would also slow the copy engine down (4 ranks: 1.51e8 packed 24 % at 18 GB/s,
                // 1.70e8 raw), so the chunk goes over raw
                if (s.pack_rate < 0.6 * s.dma_rate) f = 1.0;
                // keep both estimates alive: a tenth of every 32nd chunk goes the other way
                if (f == 1.0 && (++s.pack_probe & 31) == 0) f = 0.9;
                else if (f == 0.0 && (++s.pack_probe & 31) == 0) f = 0.1;
            } else if (h->pack_mode == 1 && s.pack_rate < 10.0e9) {
                f = 1.0;   // pageable source and one or two slow host threads: the driver's staged upload is no slower
            }
            n_raw = std::min<int64_t>(n, (int64_t)(f * (double)n + 0.5));
        }
        const int64_t n_cnt = n - n_raw;
        bool staged = false;
        // upload: the workspace must have been drained by the download of chunk k - 3
        if (k >= 3) CU_TRY(h, cudaStreamWaitEvent(s_in, w.ev_out, 0));
        if (pack) {
            // the raw part first: the copy engine works on it while the host packs the rest
            if (n_raw > 0) {
                const bool timed = !s.dma_pending && pinned_src && n_raw >= 16;
                if (timed) {
                    if (!s.dma_ev[0]) { CU_TRY(h, cudaEventCreate(&s.dma_ev[0])); CU_TRY(h, cudaEventCreate(&s.dma_ev[1])); }
                    CU_TRY(h, cudaEventRecord(s.dma_ev[0], s_in));
                }
                CU_TRY(h, cudaMemcpyAsync(w.signal, io.signal + ob * T, (size_t)n_raw * B * T * sizeof(double), cudaMemcpyHostToDevice, s_in));
                if (timed) {
                    CU_TRY(h, cudaEventRecord(s.dma_ev[1], s_in));
                    s.dma_pending = true;
                    s.dma_bytes = (double)n_raw * B * T * sizeof(double);
                }
            }
            if (n_cnt > 0) {
                const int sb = (int)(k % 3);
                const size_t off = (size_t)n_raw * B * T, cnt = (size_t)n_cnt * B * T;
                if (s.stage_busy[sb]) CU_TRY(h, cudaEventSynchronize(s.stage_ev[sb]));
                s.stage_busy[sb] = false;
                const auto t0 = std::chrono::steady_clock::now();
                staged = s.packer->pack(io.signal + ob * T + off, s.stage[sb], cnt, h->pack_lsb);
                const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                if (staged) {
                    s.packed_chunks++;
                    s.pack_seconds += dt;
                    s.pack_bytes += (double)cnt * sizeof(double);
                    if (dt > 0) s.pack_rate = 0.5 * s.pack_rate + 0.5 * (double)cnt * sizeof(double) / dt;
                    if (!w.counts) {
                        if ((rc = dev_alloc(h, s, &w.counts, (size_t)w.io_cap * B * T))) return rc;
                    }
                    CU_TRY(h, cudaMemcpyAsync(w.counts, s.stage[sb], cnt * sizeof(int16_t), cudaMemcpyHostToDevice, s_in));
                    CU_TRY(h, cudaEventRecord(s.stage_ev[sb], s_in));
                    s.stage_busy[sb] = true;
                } else {   // not on the lattice: the caller's doubles
                    s.raw_chunks++;
                    CU_TRY(h, cudaMemcpyAsync(w.signal + off, io.signal + ob * T + off, cnt * sizeof(double), cudaMemcpyHostToDevice, s_in));
                }
            }
        } else if (io.packed) {
            const size_t words = (size_t)(io.offsets[e0 + n] - io.offsets[e0]);
            if (!w.poffs) {
                if ((rc = dev_alloc(h, s, &w.poffs, (size_t)w.cap + 1))) return rc;
            }
            if (words > w.packed_cap) {   // grows to the largest chunk seen (a full event is 1104 * 112 words)
                CU_TRY(h, cudaStreamSynchronize(s_cmp));
                CU_TRY(h, cudaStreamSynchronize(s_in));
                dev_free(s, &w.packed);
                if ((rc = dev_alloc(h, s, &w.packed, words + words / 4 + 1))) return rc;
                w.packed_cap = words + words / 4 + 1;
            }
            CU_TRY(h, cudaMemcpyAsync(w.packed, io.packed + io.offsets[e0], words * sizeof(double), cudaMemcpyHostToDevice, s_in));
            CU_TRY(h, cudaMemcpyAsync(w.poffs, io.offsets + e0, (size_t)(n + 1) * sizeof(long long), cudaMemcpyHostToDevice, s_in));
        } else if (io.counts) {
            if (!w.counts) {
                if ((rc = dev_alloc(h, s, &w.counts, (size_t)w.io_cap * B * T))) return rc;
            }
            CU_TRY(h, cudaMemcpyAsync(w.counts, io.counts + ob * T, nb * T * sizeof(int16_t), cudaMemcpyHostToDevice, s_in));
        } else {
            CU_TRY(h, cudaMemcpyAsync(w.signal, io.signal + ob * T, nb * T * sizeof(double), cudaMemcpyHostToDevice, s_in));
        }
        if (!io.packed) CU_TRY(h, cudaMemcpyAsync(w.pres, io.pres + ob, nb * sizeof(int32_t), cudaMemcpyHostToDevice, s_in));
        if (io.corr) CU_TRY(h, cudaMemcpyAsync(w.corr, io.corr + e0, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s_in));
        else CU_TRY(h, cudaMemsetAsync(w.corr, 0, (size_t)n * sizeof(double), s_in));
        CU_TRY(h, cudaEventRecord(w.ev_in, s_in));
        // compute: after the upload, and not before chunk k - 2 is done (two chunks in flight)
        CU_TRY(h, cudaStreamWaitEvent(s_cmp, w.ev_in, 0));
        if (k >= 2) CU_TRY(h, cudaStreamWaitEvent(s_cmp, s.ws[(which + 1) % 3].ev_cmp, 0));
        if (io.packed) {
            unpack_kernel<<<(unsigned)std::min<int64_t>(n, 4 * s.sm_count), UNPACK_THREADS, 0, s_cmp>>>(
                w.packed, w.poffs, (long long)io.offsets[e0], n, w.signal, w.pres);
            CU_TRY(h, cudaGetLastError());
        } else if (staged) {
            const long long tot = (long long)n_cnt * B * T;
            widen_counts_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s_cmp>>>(w.counts, w.signal + (size_t)n_raw * B * T,
                                                                                  h->pack_lsb, tot);
            CU_TRY(h, cudaGetLastError());
        } else if (io.counts) {
            const long long tot = (long long)nb * T;
            widen_counts_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s_cmp>>>(w.counts, w.signal, io.lsb, tot);
            CU_TRY(h, cudaGetLastError());
        }
        rc = run_chunk(h, s, w, s_cmp, n, w.signal, w.pres, w.corr, w.wfnpulse, w.wftime, w.wfampl, w.chi2, w.timewf,
                       w.amplwf, w.status);
        if (rc) return rc;
        if (io.flat) {
            const unsigned g = (unsigned)std::min<int64_t>(n, 8 * s.sm_count);
            flat_totals_kernel<<<g, FLAT_THREADS, 0, s_cmp>>>(w.wfnpulse, n, w.ev_total);
            flat_scan_kernel<<<1, 1024, 0, s_cmp>>>(w.ev_total, n, w.ev_off);
            flat_scatter_kernel<<<g, FLAT_THREADS, 0, s_cmp>>>(w.wfnpulse, w.wftime, w.wfampl, n, w.ev_off, w.flat_t, w.flat_a, nullptr);
            CU_TRY(h, cudaGetLastError());
        }
        CU_TRY(h, cudaEventRecord(w.ev_cmp, s_cmp));
        // the pulses of chunk k - 2 go first on the download stream: its workspace is the next to be refilled, and its
        // pulse count has long reached the host (waiting for chunk k - 1 here would stop the host from running ahead)
        if (pend2.w && (rc = finish_flat())) return rc;
        // download
        CU_TRY(h, cudaStreamWaitEvent(s_out, w.ev_cmp, 0));
        if (io.wfnpulse) CU_TRY(h, cudaMemcpyAsync(io.wfnpulse + ob, w.wfnpulse, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, s_out));
        if (io.wftime) CU_TRY(h, cudaMemcpyAsync(io.wftime + ob * MAXP, w.wftime, nb * MAXP * sizeof(double), cudaMemcpyDeviceToHost, s_out));
        if (io.wfampl) CU_TRY(h, cudaMemcpyAsync(io.wfampl + ob * MAXP, w.wfampl, nb * MAXP * sizeof(double), cudaMemcpyDeviceToHost, s_out));
        if (io.chi2) CU_TRY(h, cudaMemcpyAsync(io.chi2 + ob, w.chi2, nb * sizeof(double), cudaMemcpyDeviceToHost, s_out));
        if (io.timewf) CU_TRY(h, cudaMemcpyAsync(io.timewf + ob, w.timewf, nb * sizeof(double), cudaMemcpyDeviceToHost, s_out));
        if (io.amplwf) CU_TRY(h, cudaMemcpyAsync(io.amplwf + ob, w.amplwf, nb * sizeof(double), cudaMemcpyDeviceToHost, s_out));
        if (io.status) CU_TRY(h, cudaMemcpyAsync(io.status + ob, w.status, nb * sizeof(uint8_t), cudaMemcpyDeviceToHost, s_out));
        if (io.flat) {
            CU_TRY(h, cudaMemcpyAsync(w.h_ev_off, w.ev_off, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, s_out));
            CU_TRY(h, cudaEventRecord(w.ev_tot, s_out));
            (pend.w ? pend2 : pend) = Pending{&w, e0, n};   // ev_out is recorded by finish_flat()
        } else {
            CU_TRY(h, cudaEventRecord(w.ev_out, s_out));
        }
    }
    while (pend.w)
        if ((rc = finish_flat())) return rc;
    if (io.flat && io.pulses_out) io.pulses_out[d] = pool_cur - (int64_t)((__int128)io.pool_cap * lo / io.n_total);
    CU_TRY(h, cudaStreamSynchronize(s_in));
    for (int i = 0; i < 3; i++) CU_TRY(h, cudaStreamSynchronize(s.ws[i].stream));
    CU_TRY(h, cudaStreamSynchronize(s_out));
    return fold_profile(h, s);
}

int check_handle(npswf_handle *h)
{
    if (!h) return NPSWF_ERR_ARG;
    if (h->slots.empty()) {
        h->err = "handle has no usable CUDA device";
        return NPSWF_ERR_CUDA;
    }
    return 0;
}

}  // namespace

extern "C" {

void npswf_default_config(NpsWfConfig *cfg)
{
    if (!cfg) return;
    std::memset(cfg, 0, sizeof *cfg);
    cfg->specthres = 0.02;   // T2:70
    cfg->mfthres = 1.5;      // T2:71
    cfg->trig_thres = 10;    // T2:72
    cfg->coinc_width = 20;   // T2:73
    cfg->dt = 4.0;           // T2:354
    cfg->timerefacc = 0;     // T2:81
    cfg->n_devices = 1;
    cfg->devices = nullptr;
    cfg->chunk_events = 0;
    cfg->fit_max_iter = 0;
    cfg->fit_retry_max_iter = 0;
}

int npswf_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *npswf_last_error(const npswf_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void *npswf_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    return p;
}
void npswf_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int npswf_create(const NpsWfConfig *cfg, const NpsWfCalib *cal, npswf_handle **out)
{
    if (!cfg || !cal || !out || !cal->interpX || !cal->interpY || !cal->timeref || !cal->cortime || !cal->preswf) {
        g_create_error = "npswf_create: null argument";
        return NPSWF_ERR_ARG;
    }
    npswf_handle *h = new npswf_handle;
    h->cfg = *cfg;
    h->kp.specthres = cfg->specthres; h->kp.mfthres = cfg->mfthres; h->kp.trig_thres = cfg->trig_thres;
    h->kp.dt = cfg->dt; h->kp.timerefacc = cfg->timerefacc; h->kp.coinc_width = cfg->coinc_width;
    h->kp.fit_max_iter = cfg->fit_max_iter > 0 ? cfg->fit_max_iter : 60;
    h->kp.fit_retry_max_iter = cfg->fit_retry_max_iter > 0 ? cfg->fit_retry_max_iter : 100;
    h->kp.search_fused = getenv("NPSWF_SEARCH_FUSED") ? std::max(0, std::min(3, atoi(getenv("NPSWF_SEARCH_FUSED")))) : 1;
    h->kp.fit_thread_tries = (getenv("NPSWF_FIT_THREAD_TRIES") && atoi(getenv("NPSWF_FIT_THREAD_TRIES")) > 0) ? atoi(getenv("NPSWF_FIT_THREAD_TRIES")) : 20;
    h->fit_mode = (cfg->fit_mode == NPSWF_FIT_MIGRAD || cfg->fit_mode == NPSWF_FIT_VM) ? cfg->fit_mode : NPSWF_FIT_FAST;
    if (getenv("NPSWF_FIT_MODE")) {   // A/B runs of unchanged callers
        const int m = atoi(getenv("NPSWF_FIT_MODE"));
        h->fit_mode = (m == NPSWF_FIT_MIGRAD || m == NPSWF_FIT_VM) ? m : NPSWF_FIT_FAST;
    }
    h->chunk = cfg->chunk_events > 0 ? cfg->chunk_events : 1184;
    h->chunk_fixed = cfg->chunk_events > 0;
    h->dev_cap = h->chunk_fixed ? h->chunk : 4736;
    {
        // host threads for the lossless int16 transport: the cores of this process's share of the node, at most 16
        const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
        const int local = (getenv("LOCAL_WORLD_SIZE") && atoi(getenv("LOCAL_WORLD_SIZE")) > 0) ? atoi(getenv("LOCAL_WORLD_SIZE")) : 1;
        const int nd = std::max(1, cfg->n_devices);
        h->pack_threads = std::min(16, std::max(1, hw / (local * nd)));
        if (getenv("NPSWF_HOST_PACK_THREADS") && atoi(getenv("NPSWF_HOST_PACK_THREADS")) > 0)
            h->pack_threads = atoi(getenv("NPSWF_HOST_PACK_THREADS"));
        if (getenv("NPSWF_HOST_PACK")) h->pack_mode = std::min(2, std::max(0, atoi(getenv("NPSWF_HOST_PACK"))));
        if (getenv("NPSWF_CHUNK_RAMP")) h->chunk_ramp = atoi(getenv("NPSWF_CHUNK_RAMP")) != 0;
    }
    // ---- derived calibration on the host (T2:440-451 for mfyref/mfint; spline coefficients)
    h->mfyref.assign((size_t)B * MFW, 0.0);
    h->mfint.assign(B, 0.0);
    h->spline.assign((size_t)B * (T - 1) * 4, 0.0);
    h->timeref.assign(cal->timeref, cal->timeref + B);
    std::vector<double> mfrecip(B, 0.0), mfc((size_t)B * MFW, 0.0), mfepsf(B, 0.0);
    for (int i = 0; i < B; i++) {
        if (cal->preswf[i] != 1) continue;
        const double *X = cal->interpX + (size_t)i * T, *Y = cal->interpY + (size_t)i * T;
        for (int it = 0; it < T; it++) {
            if (std::fabs(cal->timeref[i] - X[it]) < 0.001) {
                for (int jt = 0; jt < MFW; jt++) {
                    const int idx = it + jt - MFLEFT;
                    const double v = (idx >= 0 && idx < T) ? Y[idx] : 0.0;
                    h->mfyref[(size_t)i * MFW + jt] = v;
                    h->mfint[i] += v;
                }
            }
        }
        for (int it = 1; it < T; it++) {
            if (!(X[it] > X[it - 1])) {
                g_create_error = "npswf_create: interpX must be strictly increasing";
                delete h;
                return NPSWF_ERR_CALIB;
            }
        }
        {   // the fast kernels index the spline by floor(x): that needs the knots to be the sample indices 0..109.  Any
            // other strictly increasing interpX (T2:432) takes the generic-knot path (bisection per evaluation, as GSL
            // does); it must cover the model's guard interval 1 < x - t < 109 (T2:629), outside of which the reference
            // would evaluate GSL's spline out of its domain.
            bool unit = true;
            for (int it = 0; it < T; it++) unit = unit && (X[it] == (double)it);
            if (!unit) {
                h->unit_knots = false;
                if (X[0] > 1.0 || X[T - 1] < (double)(T - 1)) {
                    g_create_error = "npswf_create: interpX must cover [1, 109] (the model evaluates the spline on 1 < x - t < 109)";
                    delete h;
                    return NPSWF_ERR_CALIB;
                }
            }
        }
        if (h->mfint[i] == 0.0) {
            g_create_error = "npswf_create: matched-filter integral is zero for a block with preswf=1";
            delete h;
            return NPSWF_ERR_CALIB;
        }
        mfrecip[i] = 1.0 / h->mfint[i];
        {
            double sabs = 0.0;
            for (int jt = 0; jt < MFW; jt++) {
                mfc[(size_t)i * MFW + jt] = h->mfyref[(size_t)i * MFW + jt] * mfrecip[i];
                sabs += std::fabs(h->mfyref[(size_t)i * MFW + jt]);
            }
            mfepsf[i] = std::ldexp(sabs * std::fabs(mfrecip[i]) * (1.0 + 1e-9), -47);
        }
        build_spline(X, Y, &h->spline[(size_t)i * (T - 1) * 4]);
    }
    // ---- devices
    int ndev_avail = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev_avail);
    if (ce != cudaSuccess || ndev_avail == 0) {
        // the handle still carries the host-side derived calibration (npswf_get_spline etc.);
        // every compute entry point reports NPSWF_ERR_CUDA.  No CPU fallback.
        h->err = std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");
        (void)cudaGetLastError();
        *out = h;
        return NPSWF_OK;
    }
    int nd = cfg->n_devices > 0 ? cfg->n_devices : 1;
    for (int d = 0; d < nd; d++) h->devices.push_back(cfg->devices ? cfg->devices[d] : d);
    double ata[2 * TS_LH - 1];
    {
        const double resp[TS_LH] = {11, 43, 135, 324, 606, 882, 1000, 882, 606, 324, 135, 43, 11, 2};
        for (int lag = -(TS_LH - 1); lag <= TS_LH - 1; lag++) {
            double lda = 0;
            for (int j = 0; j < TS_LH; j++)
                if (j + lag >= 0 && j + lag < TS_LH) lda = lda + resp[j] * resp[j + lag];
            ata[lag + TS_LH - 1] = lda;
        }
    }
    h->slots.resize(nd);
    for (int d = 0; d < nd; d++) {
        DevSlot &s = h->slots[d];
        s.device = h->devices[d];
        auto fail = [&](int rc) {
            g_create_error = h->err;
            npswf_destroy(h);
            return rc;
        };
#define CR(expr)                                                          \
    do {                                                                  \
        cudaError_t _e = (expr);                                          \
        if (_e != cudaSuccess) {                                          \
            h->err = std::string(#expr ": ") + cudaGetErrorString(_e);    \
            return fail(NPSWF_ERR_CUDA);                                  \
        }                                                                 \
    } while (0)
        CR(cudaSetDevice(s.device));
        cudaDeviceProp prop;
        CR(cudaGetDeviceProperties(&prop, s.device));
        if (prop.major < 10) {
            h->err = "device is not sm_100 (Blackwell B200) class";
            return fail(NPSWF_ERR_CUDA);
        }
        s.sm_count = prop.multiProcessorCount;
        s.local_cpus = device_local_cpus(s.device);
        int rc;
        if ((rc = dev_upload(h, s, &s.cal.mfyref, h->mfyref.data(), h->mfyref.size()))) return fail(rc);
        if ((rc = dev_upload(h, s, &s.cal.mfint, h->mfint.data(), h->mfint.size()))) return fail(rc);
        if ((rc = dev_upload(h, s, &s.cal.mfrecip, mfrecip.data(), mfrecip.size()))) return fail(rc);
        if ((rc = dev_upload(h, s, &s.cal.mfc, mfc.data(), mfc.size()))) return fail(rc);
        {   // coincidence window per block: the reference's test |it - center| < coinc_width (T2:232, 267), bin by bin
            std::vector<int32_t> wlo(B, 1 << 20), wspan(B, 0);
            for (int b = 0; b < B; b++) {
                const double center = cal->timeref[b] + h->kp.timerefacc;
                int lo = -1, hi = -1;
                bool contiguous = true;
                for (int it = 0; it < T; it++) {
                    if (std::fabs((double)it - center) < (double)h->kp.coinc_width) {
                        if (lo < 0) lo = it;
                        else if (hi != it - 1) contiguous = false;
                        hi = it;
                    }
                }
                if (!contiguous) { h->err = "coincidence window is not an interval"; return fail(NPSWF_ERR_CALIB); }
                if (lo >= 0) { wlo[b] = lo; wspan[b] = hi - lo; }
            }
            if ((rc = dev_upload(h, s, &s.cal.win_lo, wlo.data(), wlo.size()))) return fail(rc);
            if ((rc = dev_upload(h, s, &s.cal.win_span, wspan.data(), wspan.size()))) return fail(rc);
        }
        if ((rc = dev_upload(h, s, &s.cal.mfepsf, mfepsf.data(), mfepsf.size()))) return fail(rc);
        if ((rc = dev_upload(h, s, &s.cal.timeref, cal->timeref, (size_t)B))) return fail(rc);
        if ((rc = dev_upload(h, s, &s.cal.cortime, cal->cortime, (size_t)B))) return fail(rc);
        if ((rc = dev_upload(h, s, &s.cal.preswf, cal->preswf, (size_t)B))) return fail(rc);
        if ((rc = dev_upload(h, s, &s.cal.spline, h->spline.data(), h->spline.size()))) return fail(rc);
        if (!h->unit_knots && (rc = dev_upload(h, s, &s.cal.knots_x, cal->interpX, (size_t)B * T))) return fail(rc);
        {   // knot form (y_i, c_i), zero padded: S on [i, i+1] from the two knots (kernel_fit_thread.cuh)
            std::vector<double2> kn((size_t)B * KN_LEN, make_double2(0.0, 0.0));
            for (int b = 0; b < B; b++) {
                const double *co = h->spline.data() + (size_t)b * (T - 1) * 4;
                for (int i = 0; i < T - 1; i++) kn[(size_t)b * KN_LEN + KN_LO + i] = make_double2(co[4 * i], co[4 * i + 2]);
                kn[(size_t)b * KN_LEN + KN_LO + T - 1] = make_double2(cal->interpY[(size_t)b * T + T - 1], 0.0);
            }
            if ((rc = dev_upload(h, s, &s.cal.knots, kn.data(), kn.size()))) return fail(rc);
        }
        {   // Gold iteration 1 (x = 1): den[i] = sum of the in-range At*A taps (integers: exact), and RN(1/den[i])
            std::vector<double> g1(2 * TS_S);
            for (int i = 0; i < TS_S; i++) {
                double den = 0;
                for (int j = -(TS_LH - 1); j <= TS_LH - 1; j++)
                    if (i + j >= 0 && i + j < TS_S) den = den + ata[j + TS_LH - 1] * 1.0;
                g1[i] = den;
                g1[TS_S + i] = 1.0 / den;
            }
            if ((rc = dev_upload(h, s, &s.gold1, g1.data(), g1.size()))) return fail(rc);
        }
        if ((rc = dev_alloc(h, s, &s.ctr, 1))) return fail(rc);
        CR(cudaMemset(s.ctr, 0, sizeof(DeviceCounters)));
        CR(cudaStreamCreateWithFlags(&s.own_stream, cudaStreamNonBlocking));
        CR(cudaStreamCreateWithFlags(&s.copy_in, cudaStreamNonBlocking));
        CR(cudaStreamCreateWithFlags(&s.copy_out, cudaStreamNonBlocking));
        s.fit_concurrent = !(getenv("NPSWF_FIT_CONCURRENT") && atoi(getenv("NPSWF_FIT_CONCURRENT")) == 0);
        CR(cudaEventCreateWithFlags(&s.fit_fork, cudaEventDisableTiming));
        CR(cudaEventCreateWithFlags(&s.last_use, cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) CR(cudaEventCreateWithFlags(&s.chunk_join[i], cudaEventDisableTiming));
        for (int i = 0; i < 4; i++) {
            CR(cudaStreamCreateWithFlags(&s.fit_stream[i], cudaStreamNonBlocking));
            CR(cudaEventCreateWithFlags(&s.fit_join[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 2; i++)
            if ((rc = alloc_workspace(h, s, s.ws[i], h->chunk, false))) return fail(rc);
        CR(cudaFuncSetAttribute(front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FRONT_SMEM));
        CR(cudaFuncSetAttribute(search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEARCH_SMEM));
        CR(cudaFuncSetAttribute(fit_kernel<25>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(sizeof(FitSmem<25>) * FIT_WARPS)));
        CR(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_front, front_kernel, FRONT_THREADS, FRONT_SMEM));
        CR(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_search, search_kernel, SEARCH_THREADS, SEARCH_SMEM));
        CR(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_fit_big, fit_kernel<25>, FIT_THREADS,
                                                         sizeof(FitSmem<25>) * FIT_WARPS));
        CR(cudaFuncSetAttribute(fit_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(sizeof(FitSmem<13>) * FIT_WARPS)));
        CR(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_fit_mid, fit_kernel<13>, FIT_THREADS,
                                                         sizeof(FitSmem<13>) * FIT_WARPS));
        CR(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_fit_small[1], fit_small_kernel<1, 8, FS_MINB1>, FS_THREADS, 0));
        CR(migrad_setup(s.occ_migrad));
        CR(migrad_thread_setup(s.occ_migrad_thread));
        s.migrad_thread = !(getenv("NPSWF_MIGRAD_THREAD") && atoi(getenv("NPSWF_MIGRAD_THREAD")) == 0);
        if ((rc = dev_alloc(h, s, &s.mg_wtab, (size_t)MIGRAD_WTAB_ENTRIES))) return fail(rc);
        if (s.occ_migrad[0] < 1 || s.occ_migrad[1] < 1 || s.occ_migrad[2] < 1) { h->err = "fit_migrad_kernel does not fit on this device"; return fail(NPSWF_ERR_CUDA); }
        s.fit_thread = !(getenv("NPSWF_FIT_THREAD") && atoi(getenv("NPSWF_FIT_THREAD")) == 0);
        CR(cudaFuncSetAttribute(fit_thread_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ft_smem(1)));
        CR(cudaFuncSetAttribute(fit_thread_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ft_smem(2)));
        CR(cudaFuncSetAttribute(fit_thread_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ft_smem(3)));
        CR(cudaFuncSetAttribute(fit_thread_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ft_smem(4)));
        CR(cudaFuncSetAttribute(fit_thread_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ft_smem(5)));
        CR(cudaFuncSetAttribute(fit_thread_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ft_smem(6)));
        CR(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_fit_thread[3], fit_thread_kernel<3>, FT_THREADS, ft_smem(3)));
#define NPSWF_VM_SETUP(NN)                                                                                                   \
    CR(cudaFuncSetAttribute(fit_vm_thread_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VM_SMEM));            \
    CR(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_vm_thread[NN], fit_vm_thread_kernel<NN>, VM_THREADS, VM_SMEM))
        NPSWF_VM_SETUP(1); NPSWF_VM_SETUP(2); NPSWF_VM_SETUP(3);
#if NPSWF_VM_MAXN >= 6
        NPSWF_VM_SETUP(4); NPSWF_VM_SETUP(5); NPSWF_VM_SETUP(6);
#else
        s.occ_vm_thread[4] = s.occ_vm_thread[5] = s.occ_vm_thread[6] = 0;
#endif
#undef NPSWF_VM_SETUP
        CR(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_fit_thread[1], fit_thread_kernel<1>, FT_THREADS, ft_smem(1)));
        CR(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_fit_thread[2], fit_thread_kernel<2>, FT_THREADS, ft_smem(2)));
        if (getenv("NPSWF_FIT_THREAD_OCC")) s.fit_thread_maxocc = atoi(getenv("NPSWF_FIT_THREAD_OCC"));
        if (s.fit_thread_maxocc > 0) {
            for (int n = 1; n <= 3; n++) s.occ_fit_thread[n] = std::min(s.occ_fit_thread[n], s.fit_thread_maxocc);
            const int carve = (int)((s.fit_thread_maxocc * (FT_SMEM + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
            CR(cudaFuncSetAttribute(fit_thread_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, std::min(100, carve)));
            CR(cudaFuncSetAttribute(fit_thread_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, std::min(100, carve)));
        }
        if (getenv("NPSWF_VERBOSE"))
            fprintf(stderr, "npswf: resident CTAs per SM: front %d search %d fit_thread<1> %d fit_thread<2> %d fit_small %d/%d/%d fit<25> %d\n",
                    s.occ_front, s.occ_search, s.occ_fit_thread[1], s.occ_fit_thread[2], s.occ_fit_small[1], s.occ_fit_small[2],
                    s.occ_fit_small[3], s.occ_fit_big);
        if (s.occ_fit_thread[1] < 1 || s.occ_fit_thread[2] < 1 || s.occ_fit_thread[3] < 1) { h->err = "fit_thread_kernel does not fit on this device"; return fail(NPSWF_ERR_CUDA); }
        CR(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_fit_small[2], fit_small_kernel<2, 8, FS_MINB2>, FS_THREADS, 0));
        CR(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_fit_small[3], fit_small_kernel<3, 16, FS_MINB3>, FS_THREADS, 0));
        if (s.occ_front < 1 || s.occ_search < 1 || s.occ_fit_big < 1 || s.occ_fit_mid < 1 || s.occ_fit_small[1] < 1 ||
            s.occ_fit_small[2] < 1 || s.occ_fit_small[3] < 1) {
            h->err = "a kernel does not fit on this device (occupancy 0)";
            return fail(NPSWF_ERR_CUDA);
        }
#undef CR
    }
    *out = h;
    return NPSWF_OK;
}

void npswf_destroy(npswf_handle *h)
{
    if (!h) return;
    for (DevSlot &s : h->slots) {
        cudaSetDevice(s.device);
        cudaDeviceSynchronize();
        for (int i = 0; i < 3; i++)
            if (s.ws[i].stream) cudaStreamDestroy(s.ws[i].stream);
        if (s.own_stream) cudaStreamDestroy(s.own_stream);
        if (s.copy_in) cudaStreamDestroy(s.copy_in);
        if (s.copy_out) cudaStreamDestroy(s.copy_out);
        for (int i = 0; i < 3; i++) {
            if (s.ws[i].ev_in) cudaEventDestroy(s.ws[i].ev_in);
            if (s.ws[i].ev_cmp) cudaEventDestroy(s.ws[i].ev_cmp);
            if (s.ws[i].ev_out) cudaEventDestroy(s.ws[i].ev_out);
            if (s.ws[i].ev_tot) cudaEventDestroy(s.ws[i].ev_tot);
            if (s.ws[i].h_ev_off) cudaFreeHost(s.ws[i].h_ev_off);
        }
        for (int i = 0; i < 4; i++) {
            if (s.fit_stream[i]) cudaStreamDestroy(s.fit_stream[i]);
            if (s.fit_join[i]) cudaEventDestroy(s.fit_join[i]);
        }
        if (s.fit_fork) cudaEventDestroy(s.fit_fork);
        if (s.last_use) cudaEventDestroy(s.last_use);
        for (int i = 0; i < 2; i++)
            if (s.chunk_join[i]) cudaEventDestroy(s.chunk_join[i]);
        for (cudaEvent_t e : s.prof_events) cudaEventDestroy(e);
        for (cudaEvent_t e : s.prof_pool) cudaEventDestroy(e);
        for (void *p : s.owned) cudaFree(p);
        for (int i = 0; i < 3; i++) {
            if (s.stage[i]) cudaFreeHost(s.stage[i]);
            if (s.stage_ev[i]) cudaEventDestroy(s.stage_ev[i]);
        }
        for (int i = 0; i < 2; i++)
            if (s.dma_ev[i]) cudaEventDestroy(s.dma_ev[i]);
    }
    delete h;
}

int npswf_set_host_packing(npswf_handle *h, int mode, int n_threads, double lsb_mV)
{
    if (!h || mode < 0 || mode > 2 || n_threads < 0 || !(lsb_mV >= 0)) return NPSWF_ERR_ARG;
    h->pack_mode = mode;
    if (lsb_mV > 0) h->pack_lsb = lsb_mV;
    if (n_threads > 0 && n_threads != h->pack_threads) {
        h->pack_threads = n_threads;
        for (DevSlot &s : h->slots) s.packer.reset();
    }
    return 0;
}

int npswf_debug_pack_counts(const double *x, int64_t n, double lsb_mV, int32_t n_threads, int16_t *counts_out)
{
    if (!x || !counts_out || n < 0 || !(lsb_mV > 0) || n_threads < 1) return NPSWF_ERR_ARG;
    PackPool pool(n_threads);
    return pool.pack(x, counts_out, (size_t)n, lsb_mV) ? 1 : 0;
}

int npswf_host_packing_stats(const npswf_handle *h, int64_t *packed_chunks, int64_t *raw_chunks, double *pack_gb_per_s,
                             int64_t *packed_input_bytes)
{
    if (!h) return NPSWF_ERR_ARG;
    int64_t p = 0, r = 0;
    double sec = 0, bytes = 0;
    for (const DevSlot &s : h->slots) { p += s.packed_chunks; r += s.raw_chunks; sec += s.pack_seconds; bytes += s.pack_bytes; }
    if (packed_chunks) *packed_chunks = p;
    if (raw_chunks) *raw_chunks = r;
    if (pack_gb_per_s) *pack_gb_per_s = sec > 0 ? bytes / sec / 1e9 : 0.0;
    if (packed_input_bytes) *packed_input_bytes = (int64_t)bytes;
    return 0;
}

int npswf_host_upload_rate(const npswf_handle *h, double *raw_gb_per_s, int32_t *numa_bound_cpus)
{
    if (!h) return NPSWF_ERR_ARG;
    double r = 0;
    int cpus = 0;
    for (const DevSlot &s : h->slots) { r += s.dma_rate; cpus += (int)s.local_cpus.size(); }
    if (raw_gb_per_s) *raw_gb_per_s = h->slots.empty() ? 0.0 : r / (double)h->slots.size() / 1e9;
    if (numa_bound_cpus) *numa_bound_cpus = cpus;
    return 0;
}

int npswf_get_counters(npswf_handle *h, NpsWfCounters *out)
{
    if (!h || !out) return NPSWF_ERR_ARG;
    NpsWfCounters c = h->host_ctr;
    for (DevSlot &s : h->slots) {
        DeviceCounters dc;
        CU_TRY(h, cudaSetDevice(s.device));
        CU_TRY(h, cudaDeviceSynchronize());
        CU_TRY(h, cudaMemcpy(&dc, s.ctr, sizeof dc, cudaMemcpyDeviceToHost));
        c.n_present += (int64_t)dc.n_present;
        c.n_pass_threshold += (int64_t)dc.n_pass_threshold;
        c.n_fit_attempted += (int64_t)dc.n_fit_attempted;
        c.n_fit_ok_first += (int64_t)dc.n_fit_ok_first;
        c.n_fit_ok_retry += (int64_t)dc.n_fit_ok_retry;
        c.n_fallback += (int64_t)dc.n_fallback;
        c.n_pulses += (int64_t)dc.n_pulses;
        c.n_peak_buffer_full += (int64_t)dc.n_peak_buffer_full;
        c.n_fit_iterations += (int64_t)dc.n_fit_iterations;
        c.n_fit_evals += (int64_t)dc.n_fit_evals;
    }
    *out = c;
    return 0;
}

int npswf_reset_counters(npswf_handle *h)
{
    if (!h) return NPSWF_ERR_ARG;
    h->host_ctr = NpsWfCounters{};
    for (DevSlot &s : h->slots) {
        CU_TRY(h, cudaSetDevice(s.device));
        CU_TRY(h, cudaDeviceSynchronize());
        CU_TRY(h, cudaMemset(s.ctr, 0, sizeof(DeviceCounters)));
    }
    return 0;
}

int npswf_analyze_batch(npswf_handle *h, int64_t n_events, const double *signal, const int32_t *pres,
                        const double *corr_time_HMS, int32_t *wfnpulse, double *wftime, double *wfampl, double *chi2,
                        double *timewf, double *amplwf, uint8_t *status)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_events < 0 || (n_events > 0 && (!signal || !pres))) {
        h->err = "npswf_analyze_batch: bad arguments";
        return NPSWF_ERR_ARG;
    }
    if (n_events == 0) return 0;
    HostIO io;
    io.signal = signal; io.pres = pres; io.corr = corr_time_HMS; io.wfnpulse = wfnpulse; io.wftime = wftime;
    io.wfampl = wfampl; io.chi2 = chi2; io.timewf = timewf; io.amplwf = amplwf; io.status = status;
    rc = for_each_slot_range(h, n_events, [&](int d, int64_t lo, int64_t hi) { return analyze_range(h, d, lo, hi, io); });
    if (rc) return rc;
    h->host_ctr.n_events += n_events;
    h->host_ctr.n_block_waveforms += n_events * B;
    return 0;
}

int npswf_analyze_batch_flat(npswf_handle *h, int64_t n_events, const double *signal, const int32_t *pres,
                             const double *corr_time_HMS, int32_t *wfnpulse, int64_t *pulse_offset, int32_t *pulse_count,
                             double *wftime_pool, double *wfampl_pool, int64_t pool_capacity, double *chi2, double *timewf,
                             double *amplwf, uint8_t *status, int64_t *n_pulses)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_events < 0 || pool_capacity < 0 ||
        (n_events > 0 && (!signal || !pres || !pulse_offset || !pulse_count || !wftime_pool || !wfampl_pool))) {
        h->err = "npswf_analyze_batch_flat: bad arguments";
        return NPSWF_ERR_ARG;
    }
    if (n_pulses) *n_pulses = 0;
    if (n_events == 0) return 0;
    std::vector<int64_t> per_dev(h->slots.size(), 0);
    HostIO io;
    io.signal = signal; io.pres = pres; io.corr = corr_time_HMS; io.wfnpulse = wfnpulse; io.chi2 = chi2; io.timewf = timewf;
    io.amplwf = amplwf; io.status = status;
    io.flat = true; io.pulse_offset = pulse_offset; io.pulse_count = pulse_count; io.pool_t = wftime_pool; io.pool_a = wfampl_pool;
    io.pool_cap = pool_capacity; io.n_total = n_events; io.pulses_out = per_dev.data();
    rc = for_each_slot_range(h, n_events, [&](int d, int64_t lo, int64_t hi) { return analyze_range(h, d, lo, hi, io); });
    if (rc) return rc;
    if (n_pulses)
        for (int64_t v : per_dev) *n_pulses += v;
    h->host_ctr.n_events += n_events;
    h->host_ctr.n_block_waveforms += n_events * B;
    return 0;
}

int npswf_analyze_batch_flat_i16(npswf_handle *h, int64_t n_events, const int16_t *counts, double lsb_mV, const int32_t *pres,
                                 const double *corr_time_HMS, int32_t *wfnpulse, int64_t *pulse_offset, int32_t *pulse_count,
                                 double *wftime_pool, double *wfampl_pool, int64_t pool_capacity, double *chi2, double *timewf,
                                 double *amplwf, uint8_t *status, int64_t *n_pulses)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_events < 0 || pool_capacity < 0 || !(lsb_mV > 0) ||
        (n_events > 0 && (!counts || !pres || !pulse_offset || !pulse_count || !wftime_pool || !wfampl_pool))) {
        h->err = "npswf_analyze_batch_flat_i16: bad arguments";
        return NPSWF_ERR_ARG;
    }
    if (n_pulses) *n_pulses = 0;
    if (n_events == 0) return 0;
    std::vector<int64_t> per_dev(h->slots.size(), 0);
    HostIO io;
    io.counts = counts; io.lsb = lsb_mV; io.pres = pres; io.corr = corr_time_HMS; io.wfnpulse = wfnpulse; io.chi2 = chi2;
    io.timewf = timewf; io.amplwf = amplwf; io.status = status;
    io.flat = true; io.pulse_offset = pulse_offset; io.pulse_count = pulse_count; io.pool_t = wftime_pool; io.pool_a = wfampl_pool;
    io.pool_cap = pool_capacity; io.n_total = n_events; io.pulses_out = per_dev.data();
    rc = for_each_slot_range(h, n_events, [&](int d, int64_t lo, int64_t hi) { return analyze_range(h, d, lo, hi, io); });
    if (rc) return rc;
    if (n_pulses)
        for (int64_t v : per_dev) *n_pulses += v;
    h->host_ctr.n_events += n_events;
    h->host_ctr.n_block_waveforms += n_events * B;
    return 0;
}

int npswf_analyze_batch_i16(npswf_handle *h, int64_t n_events, const int16_t *counts, double lsb_mV,
                            const int32_t *pres, const double *corr_time_HMS, int32_t *wfnpulse, double *wftime,
                            double *wfampl, double *chi2, double *timewf, double *amplwf, uint8_t *status)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_events < 0 || (n_events > 0 && (!counts || !pres))) {
        h->err = "npswf_analyze_batch_i16: bad arguments";
        return NPSWF_ERR_ARG;
    }
    if (n_events == 0) return 0;
    HostIO io;
    io.counts = counts; io.lsb = lsb_mV; io.pres = pres; io.corr = corr_time_HMS; io.wfnpulse = wfnpulse;
    io.wftime = wftime; io.wfampl = wfampl; io.chi2 = chi2; io.timewf = timewf; io.amplwf = amplwf; io.status = status;
    rc = for_each_slot_range(h, n_events, [&](int d, int64_t lo, int64_t hi) { return analyze_range(h, d, lo, hi, io); });
    if (rc) return rc;
    h->host_ctr.n_events += n_events;
    h->host_ctr.n_block_waveforms += n_events * B;
    return 0;
}

int npswf_analyze_batch_packed(npswf_handle *h, int64_t n_events, const double *samp, const int64_t *offsets,
                               const double *corr_time_HMS, int32_t *wfnpulse, double *wftime, double *wfampl, double *chi2,
                               double *timewf, double *amplwf, uint8_t *status)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_events < 0 || (n_events > 0 && (!samp || !offsets))) {
        h->err = "npswf_analyze_batch_packed: bad arguments";
        return NPSWF_ERR_ARG;
    }
    if (n_events == 0) return 0;
    HostIO io;
    io.packed = samp; io.offsets = offsets; io.corr = corr_time_HMS; io.wfnpulse = wfnpulse;
    io.wftime = wftime; io.wfampl = wfampl; io.chi2 = chi2; io.timewf = timewf; io.amplwf = amplwf; io.status = status;
    rc = for_each_slot_range(h, n_events, [&](int d, int64_t lo, int64_t hi) { return analyze_range(h, d, lo, hi, io); });
    if (rc) return rc;
    h->host_ctr.n_events += n_events;
    h->host_ctr.n_block_waveforms += n_events * B;
    return 0;
}

int npswf_unpack_batch(npswf_handle *h, int64_t n_events, const double *samp, const int64_t *offsets, double *signal,
                       int32_t *pres)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_events < 0 || (n_events > 0 && (!samp || !offsets || !signal || !pres))) return NPSWF_ERR_ARG;
    if (n_events == 0) return 0;
    DevSlot &s = h->slots[0];
    CU_TRY(h, cudaSetDevice(s.device));
    if ((rc = wait_last_use(h, s))) return rc;
    const size_t words = (size_t)(offsets[n_events] - offsets[0]);
    DevBuf<double> d_s, d_sig;
    DevBuf<long long> d_o;
    DevBuf<int32_t> d_p;
    CU_TRY(h, d_s.alloc(words));
    CU_TRY(h, d_o.alloc((size_t)(n_events + 1)));
    CU_TRY(h, d_sig.alloc((size_t)n_events * B * T));
    CU_TRY(h, d_p.alloc((size_t)n_events * B));
    CU_TRY(h, cudaMemcpy(d_s, samp + offsets[0], words * sizeof(double), cudaMemcpyHostToDevice));
    CU_TRY(h, cudaMemcpy(d_o, offsets, (size_t)(n_events + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    unpack_kernel<<<(unsigned)std::min<int64_t>(n_events, 4 * s.sm_count), UNPACK_THREADS>>>(d_s, d_o, (long long)offsets[0],
                                                                                              n_events, d_sig, d_p);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpy(signal, d_sig, (size_t)n_events * B * T * sizeof(double), cudaMemcpyDeviceToHost));
    CU_TRY(h, cudaMemcpy(pres, d_p, (size_t)n_events * B * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

int npswf_event_diagnostics_batch(npswf_handle *h, int64_t n_events, const double *signal, double *ampl, double *enertot,
                                  double *integtot)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_events < 0 || (n_events > 0 && !signal)) return NPSWF_ERR_ARG;
    if (n_events == 0) return 0;
    DevSlot &s = h->slots[0];
    CU_TRY(h, cudaSetDevice(s.device));
    if ((rc = wait_last_use(h, s))) return rc;
    DevBuf<double> d_sig, d_a, d_e, d_i;
    CU_TRY(h, d_sig.alloc((size_t)n_events * B * T));
    CU_TRY(h, d_a.alloc((size_t)n_events * B));
    CU_TRY(h, d_e.alloc((size_t)n_events));
    CU_TRY(h, d_i.alloc((size_t)n_events));
    CU_TRY(h, cudaMemcpy(d_sig, signal, (size_t)n_events * B * T * sizeof(double), cudaMemcpyHostToDevice));
    diag_kernel<<<(unsigned)std::min<int64_t>(n_events, 8 * s.sm_count), DIAG_THREADS>>>(d_sig, n_events, d_a, d_e, d_i);
    CU_TRY(h, cudaGetLastError());
    if (ampl) CU_TRY(h, cudaMemcpy(ampl, d_a, (size_t)n_events * B * sizeof(double), cudaMemcpyDeviceToHost));
    if (enertot) CU_TRY(h, cudaMemcpy(enertot, d_e, (size_t)n_events * sizeof(double), cudaMemcpyDeviceToHost));
    if (integtot) CU_TRY(h, cudaMemcpy(integtot, d_i, (size_t)n_events * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int npswf_event_diagnostics_device(npswf_handle *h, int32_t dev_slot, int64_t n_events, const double *d_signal, double *d_ampl,
                                   double *d_enertot, double *d_integtot, void *stream)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (dev_slot < 0 || dev_slot >= (int)h->slots.size() || n_events < 0 || !d_signal) return NPSWF_ERR_ARG;
    if (n_events == 0) return 0;
    DevSlot &s = h->slots[dev_slot];
    CU_TRY(h, cudaSetDevice(s.device));
    cudaStream_t st = stream ? (cudaStream_t)stream : s.own_stream;
    diag_kernel<<<(unsigned)std::min<int64_t>(n_events, 8 * s.sm_count), DIAG_THREADS, 0, st>>>(d_signal, n_events, d_ampl,
                                                                                               d_enertot, d_integtot);
    CU_TRY(h, cudaGetLastError());
    return 0;
}

int npswf_analyze_batch_device(npswf_handle *h, int32_t dev_slot, int64_t n_events, const double *d_signal,
                               const int32_t *d_pres, const double *d_corr, int32_t *d_wfnpulse, double *d_wftime,
                               double *d_wfampl, double *d_chi2, double *d_timewf, double *d_amplwf, uint8_t *d_status,
                               void *stream)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (dev_slot < 0 || dev_slot >= (int)h->slots.size() || n_events < 0 || !d_signal || !d_pres || !d_wftime ||
        !d_wfampl || !d_chi2 || !d_status) {
        h->err = "npswf_analyze_batch_device: bad arguments (wftime/wfampl/chi2/status are required)";
        return NPSWF_ERR_ARG;
    }
    if (((uintptr_t)d_signal & 15) != 0) {
        h->err = "npswf_analyze_batch_device: d_signal must be 16-byte aligned (bulk TMA source)";
        return NPSWF_ERR_ARG;
    }
    DevSlot &s = h->slots[dev_slot];
    CU_TRY(h, cudaSetDevice(s.device));
    cudaStream_t st = stream ? (cudaStream_t)stream : s.own_stream;
    // an earlier device-path call (possibly on another stream) may still be using the shared scratch and fit streams
    if (s.last_use_valid) CU_TRY(h, cudaStreamWaitEvent(st, s.last_use, 0));
    // Chunks alternate between the two workspaces, each on its own internal stream forked from / joined to the
    // caller's stream: the front + search kernels of chunk k+1 fill the SMs the fit tails of chunk k leave idle.
    // With stage profiling on everything is serialised on the caller's stream so that the stage times are clean.
    int64_t chunk = h->chunk;
    if (!h->chunk_fixed) {
        chunk = std::min<int64_t>(h->dev_cap, std::max<int64_t>(h->chunk, ((n_events + 1) / 2 + 147) / 148 * 148));
        if ((rc = grow_scratch(h, s, chunk))) return rc;
    }
    // (the Migrad kernels are long-running persistent grids that fill the device on their own: two 4 736-event chunks'
    // worth of them co-scheduled run 20-50 % slower than one after the other, so that mode keeps its chunks in
    // sequence here; in the host pipeline, whose chunks are a quarter of that, the overlap still pays: 8.9 vs 7.1 M/s)
    const bool overlap = !h->profiling && n_events > chunk && h->fit_mode != NPSWF_FIT_MIGRAD;
    if (overlap) {
        CU_TRY(h, cudaEventRecord(s.fit_fork, st));
        for (int i = 0; i < 2; i++) CU_TRY(h, cudaStreamWaitEvent(s.ws[i].stream, s.fit_fork, 0));
    }
    int which = 0;
    for (int64_t e0 = 0; e0 < n_events; e0 += chunk, which ^= 1) {
        Workspace &w = s.ws[overlap ? which : 0];
        const int64_t n = std::min<int64_t>(chunk, n_events - e0);
        const size_t ob = (size_t)e0 * B;
        rc = run_chunk(h, s, w, overlap ? w.stream : st, n, d_signal + ob * T, d_pres + ob, d_corr ? d_corr + e0 : nullptr,
                       d_wfnpulse ? d_wfnpulse + ob : nullptr, d_wftime + ob * MAXP, d_wfampl + ob * MAXP, d_chi2 + ob,
                       d_timewf ? d_timewf + ob : nullptr, d_amplwf ? d_amplwf + ob : nullptr, d_status + ob);
        if (rc) return rc;
    }
    if (overlap) {
        for (int i = 0; i < 2; i++) {
            CU_TRY(h, cudaEventRecord(s.chunk_join[i], s.ws[i].stream));
            CU_TRY(h, cudaStreamWaitEvent(st, s.chunk_join[i], 0));
        }
    }
    CU_TRY(h, cudaEventRecord(s.last_use, st));
    s.last_use_valid = true;
    h->host_ctr.n_events += n_events;
    h->host_ctr.n_block_waveforms += n_events * B;
    return 0;
}

int npswf_sync_device(npswf_handle *h, int32_t dev_slot, void *stream)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (dev_slot < 0 || dev_slot >= (int)h->slots.size()) return NPSWF_ERR_ARG;
    DevSlot &s = h->slots[dev_slot];
    CU_TRY(h, cudaSetDevice(s.device));
    CU_TRY(h, cudaStreamSynchronize(stream ? (cudaStream_t)stream : s.own_stream));
    return fold_profile(h, s);
}

int npswf_set_profiling(npswf_handle *h, int on)
{
    if (!h) return NPSWF_ERR_ARG;
    h->profiling = on != 0;
    return 0;
}

int npswf_get_stage_times(npswf_handle *h, double *ms_front, double *ms_search, double *ms_fit, int64_t *n_chunks,
                          int reset)
{
    if (!h) return NPSWF_ERR_ARG;
    if (ms_front) *ms_front = h->stage_ms[0];
    if (ms_search) *ms_search = h->stage_ms[1];
    if (ms_fit) *ms_fit = h->stage_ms[2];
    if (n_chunks) *n_chunks = h->stage_chunks;
    if (reset) { h->stage_ms[0] = h->stage_ms[1] = h->stage_ms[2] = 0; h->stage_chunks = 0; }
    return 0;
}

// ---- stage-level entry points: single device (slot 0), chunked, synchronous ----

int npswf_find_pulses_mf_batch(npswf_handle *h, int64_t n_events, const double *signal, const int32_t *pres,
                               int32_t *wfnpulse, double *wftime, double *wfampl)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_events < 0 || (n_events > 0 && (!signal || !pres || !wfnpulse || !wftime || !wfampl))) return NPSWF_ERR_ARG;
    DevSlot &s = h->slots[0];
    CU_TRY(h, cudaSetDevice(s.device));
    if ((rc = wait_last_use(h, s))) return rc;
    if ((rc = ensure_io(h, s))) return rc;
    Workspace &w = s.ws[0];
    cudaStream_t st = w.stream;
    for (int64_t e0 = 0; e0 < n_events; e0 += w.io_cap) {
        const int64_t n = std::min<int64_t>(w.io_cap, n_events - e0);
        const size_t nb = (size_t)n * B, ob = (size_t)e0 * B;
        CU_TRY(h, cudaMemcpyAsync(w.signal, signal + ob * T, nb * T * sizeof(double), cudaMemcpyHostToDevice, st));
        CU_TRY(h, cudaMemcpyAsync(w.pres, pres + ob, nb * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        if ((rc = launch_front(h, s, st, w.signal, w.pres, n, w.mf, w.minsig, w.flags, 1, 0))) return rc;
        SearchArgs sa{};
        sa.hist = w.mf; sa.flags = w.flags; sa.minsig = w.minsig; sa.signal = w.signal; sa.n_items = (long long)nb;
        sa.wfnpulse = w.wfnpulse; sa.wftime = w.wftime; sa.wfampl = w.wfampl;
        if ((rc = launch_search(h, s, st, sa))) return rc;
        CU_TRY(h, cudaMemcpyAsync(wfnpulse + ob, w.wfnpulse, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaMemcpyAsync(wftime + ob * MAXP, w.wftime, nb * MAXP * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaMemcpyAsync(wfampl + ob * MAXP, w.wfampl, nb * MAXP * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaStreamSynchronize(st));
    }
    return 0;
}

int npswf_pass_cluster_threshold_batch(npswf_handle *h, int64_t n_events, const double *signal, const int32_t *pres,
                                       uint8_t *ok)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_events < 0 || (n_events > 0 && (!signal || !pres || !ok))) return NPSWF_ERR_ARG;
    DevSlot &s = h->slots[0];
    CU_TRY(h, cudaSetDevice(s.device));
    if ((rc = wait_last_use(h, s))) return rc;
    if ((rc = ensure_io(h, s))) return rc;
    Workspace &w = s.ws[0];
    cudaStream_t st = w.stream;
    std::vector<uint8_t> tmp;
    for (int64_t e0 = 0; e0 < n_events; e0 += w.io_cap) {
        const int64_t n = std::min<int64_t>(w.io_cap, n_events - e0);
        const size_t nb = (size_t)n * B, ob = (size_t)e0 * B;
        CU_TRY(h, cudaMemcpyAsync(w.signal, signal + ob * T, nb * T * sizeof(double), cudaMemcpyHostToDevice, st));
        CU_TRY(h, cudaMemcpyAsync(w.pres, pres + ob, nb * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        if ((rc = launch_front(h, s, st, w.signal, w.pres, n, nullptr, nullptr, w.flags, 0, 1))) return rc;
        tmp.resize(nb);
        CU_TRY(h, cudaMemcpyAsync(tmp.data(), w.flags, nb, cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaStreamSynchronize(st));
        for (size_t i = 0; i < nb; i++) ok[ob + i] = (tmp[i] & FL_OKTOFIT) ? 1 : 0;
    }
    return 0;
}

int npswf_matched_filter_batch(npswf_handle *h, int64_t n_events, const double *signal, const int32_t *pres, float *mf)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_events < 0 || (n_events > 0 && (!signal || !pres || !mf))) return NPSWF_ERR_ARG;
    DevSlot &s = h->slots[0];
    CU_TRY(h, cudaSetDevice(s.device));
    if ((rc = wait_last_use(h, s))) return rc;
    if ((rc = ensure_io(h, s))) return rc;
    Workspace &w = s.ws[0];
    cudaStream_t st = w.stream;
    for (int64_t e0 = 0; e0 < n_events; e0 += w.io_cap) {
        const int64_t n = std::min<int64_t>(w.io_cap, n_events - e0);
        const size_t nb = (size_t)n * B, ob = (size_t)e0 * B;
        CU_TRY(h, cudaMemcpyAsync(w.signal, signal + ob * T, nb * T * sizeof(double), cudaMemcpyHostToDevice, st));
        CU_TRY(h, cudaMemcpyAsync(w.pres, pres + ob, nb * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CU_TRY(h, cudaMemsetAsync(w.mf, 0, nb * T * sizeof(float), st));
        if ((rc = launch_front(h, s, st, w.signal, w.pres, n, w.mf, w.minsig, w.flags, 1, 0))) return rc;
        CU_TRY(h, cudaMemcpyAsync(mf + ob * T, w.mf, nb * T * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaStreamSynchronize(st));
    }
    return 0;
}

int npswf_fitwf_batch(npswf_handle *h, int64_t n_events, const double *signal, const double *corr_time_HMS,
                      const uint8_t *fit_mask, const int32_t *wfnpulse, double *wftime, double *wfampl, double *chi2,
                      uint8_t *status)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_events < 0 || (n_events > 0 && (!signal || !fit_mask || !wfnpulse || !wftime || !wfampl || !chi2 || !status)))
        return NPSWF_ERR_ARG;
    DevSlot &s = h->slots[0];
    CU_TRY(h, cudaSetDevice(s.device));
    if ((rc = wait_last_use(h, s))) return rc;
    if ((rc = ensure_io(h, s))) return rc;
    Workspace &w = s.ws[0];
    cudaStream_t st = w.stream;
    for (int64_t e0 = 0; e0 < n_events; e0 += w.io_cap) {
        const int64_t n = std::min<int64_t>(w.io_cap, n_events - e0);
        const size_t nb = (size_t)n * B, ob = (size_t)e0 * B;
        CU_TRY(h, cudaMemcpyAsync(w.signal, signal + ob * T, nb * T * sizeof(double), cudaMemcpyHostToDevice, st));
        if (corr_time_HMS) CU_TRY(h, cudaMemcpyAsync(w.corr, corr_time_HMS + e0, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
        else CU_TRY(h, cudaMemsetAsync(w.corr, 0, (size_t)n * sizeof(double), st));
        CU_TRY(h, cudaMemcpyAsync(w.mask, fit_mask + ob, nb, cudaMemcpyHostToDevice, st));
        CU_TRY(h, cudaMemcpyAsync(w.wfnpulse, wfnpulse + ob, nb * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CU_TRY(h, cudaMemcpyAsync(w.wftime, wftime + ob * MAXP, nb * MAXP * sizeof(double), cudaMemcpyHostToDevice, st));
        CU_TRY(h, cudaMemcpyAsync(w.wfampl, wfampl + ob * MAXP, nb * MAXP * sizeof(double), cudaMemcpyHostToDevice, st));
        CU_TRY(h, cudaMemsetAsync(w.fit_count, 0, 128 * sizeof(int), st));
        CU_TRY(h, cudaMemsetAsync(w.bucket_count, 0, (size_t)(MAXP + 1) * B * sizeof(int), st));
        build_jobs_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(w.mask, w.wfnpulse, (long long)nb, w.bucket_count,
                                                                       w.fit_list, (int)w.cap, w.chi2, w.status);
        CU_TRY(h, cudaGetLastError());
        if ((rc = launch_compact(h, st, w))) return rc;
        CU_TRY(h, cudaGetLastError());
        if ((rc = launch_fits(h, s, st, w, w.signal, w.corr, w.wftime, w.wfampl, w.chi2, nullptr, nullptr, w.status))) return rc;
        CU_TRY(h, cudaMemcpyAsync(wftime + ob * MAXP, w.wftime, nb * MAXP * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaMemcpyAsync(wfampl + ob * MAXP, w.wfampl, nb * MAXP * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaMemcpyAsync(chi2 + ob, w.chi2, nb * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaMemcpyAsync(status + ob, w.status, nb, cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaStreamSynchronize(st));
    }
    return 0;
}

int npswf_tspectrum_debug(npswf_handle *h, int64_t n, const float *hist, int32_t *npeaks, double *pos_x, double *smoothed,
                          double *decon)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n < 0 || (n > 0 && !hist)) return NPSWF_ERR_ARG;
    if (n == 0) return 0;
    DevSlot &s = h->slots[0];
    CU_TRY(h, cudaSetDevice(s.device));
    if ((rc = wait_last_use(h, s))) return rc;
    DevBuf<float> d_h;
    DevBuf<int32_t> d_n;
    DevBuf<double> d_p, d_s, d_d;
    CU_TRY(h, d_h.alloc((size_t)n * T));
    CU_TRY(h, d_n.alloc((size_t)n));
    CU_TRY(h, d_p.alloc((size_t)n * MAXP));
    CU_TRY(h, d_s.alloc((size_t)n * TS_S));
    CU_TRY(h, d_d.alloc((size_t)n * T));
    CU_TRY(h, cudaMemcpy(d_h, hist, (size_t)n * T * sizeof(float), cudaMemcpyHostToDevice));
    SearchArgs sa{};
    sa.hist = d_h; sa.n_items = n; sa.npeaks_out = d_n; sa.pos_out = d_p; sa.smoothed_out = d_s; sa.decon_out = d_d;
    if ((rc = launch_search(h, s, nullptr, sa))) return rc;
    CU_TRY(h, cudaDeviceSynchronize());
    if (npeaks) CU_TRY(h, cudaMemcpy(npeaks, d_n, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (pos_x) CU_TRY(h, cudaMemcpy(pos_x, d_p, (size_t)n * MAXP * sizeof(double), cudaMemcpyDeviceToHost));
    if (smoothed) CU_TRY(h, cudaMemcpy(smoothed, d_s, (size_t)n * TS_S * sizeof(double), cudaMemcpyDeviceToHost));
    if (decon) CU_TRY(h, cudaMemcpy(decon, d_d, (size_t)n * T * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int npswf_debug_exp(npswf_handle *h, int64_t n, const double *x, double *y)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n <= 0 || !x || !y) return NPSWF_ERR_ARG;
    DevSlot &s = h->slots[0];
    CU_TRY(h, cudaSetDevice(s.device));
    if ((rc = wait_last_use(h, s))) return rc;
    DevBuf<double> dx, dy;
    CU_TRY(h, dx.alloc((size_t)n));
    CU_TRY(h, dy.alloc((size_t)n));
    CU_TRY(h, cudaMemcpy(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice));
    det_exp_debug_kernel<<<(unsigned)((n + 255) / 256), 256>>>(dx, dy, n);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpy(y, dy, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return 0;
}

// FP64 FMA-chain microbenchmark: 8 independent chains per thread, enough resident warps to saturate the pipe
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;  // keeps the chains alive
}

int npswf_debug_fp64_peak(npswf_handle *h, double *gflops)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (!gflops) return NPSWF_ERR_ARG;
    DevSlot &s = h->slots[0];
    CU_TRY(h, cudaSetDevice(s.device));
    if ((rc = wait_last_use(h, s))) return rc;
    DevBuf<double> d;
    CU_TRY(h, d.alloc(1));
    cudaEvent_t e0, e1;
    CU_TRY(h, cudaEventCreate(&e0));
    CU_TRY(h, cudaEventCreate(&e1));
    const int blocks = s.sm_count * 8, iters = 1 << 15;
    fp64_peak_kernel<<<blocks, 256>>>(d, 1024, 0.999999, 1e-9);  // warm-up
    CU_TRY(h, cudaEventRecord(e0));
    fp64_peak_kernel<<<blocks, 256>>>(d, iters, 0.999999, 1e-9);
    CU_TRY(h, cudaEventRecord(e1));
    CU_TRY(h, cudaEventSynchronize(e1));
    float ms = 0;
    CU_TRY(h, cudaEventElapsedTime(&ms, e0, e1));
    *gflops = 2.0 * 8.0 * (double)iters * 256.0 * blocks / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

int npswf_debug_exact_ops(npswf_handle *h, int64_t n_trials, uint64_t seed, uint64_t mismatch[2])
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (n_trials <= 0 || !mismatch) return NPSWF_ERR_ARG;
    DevSlot &s = h->slots[0];
    CU_TRY(h, cudaSetDevice(s.device));
    if ((rc = wait_last_use(h, s))) return rc;
    DevBuf<unsigned long long> d;
    CU_TRY(h, d.alloc(2));
    CU_TRY(h, cudaMemset(d, 0, 16));
    const int threads = 256, blocks = s.sm_count * 8;
    const int per_thread = (int)std::min<int64_t>(1 << 20, (n_trials + (int64_t)threads * blocks - 1) / ((int64_t)threads * blocks));
    exact_ops_check_kernel<<<blocks, threads>>>((unsigned long long)seed, per_thread, d);
    CU_TRY(h, cudaGetLastError());
    unsigned long long out[2] = {0, 0};
    CU_TRY(h, cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost));
    mismatch[0] = out[0];
    mismatch[1] = out[1];
    return 0;
}

int npswf_debug_vm_reasons(npswf_handle *h, uint64_t out[8], int reset)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (!out) return NPSWF_ERR_ARG;
    for (int i = 0; i < 8; i++) out[i] = 0;
    for (size_t d = 0; d < h->slots.size(); d++) {   // the tallies live on each device the handle runs on
        CU_TRY(h, cudaSetDevice(h->slots[d].device));
        CU_TRY(h, cudaDeviceSynchronize());
        unsigned long long tmp[8];
        CU_TRY(h, cudaMemcpyFromSymbol(tmp, g_vm_reason, sizeof tmp));
        for (int i = 0; i < 8; i++) out[i] += tmp[i];
        if (reset) {
            unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            CU_TRY(h, cudaMemcpyToSymbol(g_vm_reason, z, sizeof z));
        }
    }
    return 0;
}

int npswf_debug_search_fused(npswf_handle *h, uint64_t out[2], int reset)
{
    int rc = check_handle(h);
    if (rc) return rc;
    if (!out) return NPSWF_ERR_ARG;
    out[0] = out[1] = 0;
    for (size_t d = 0; d < h->slots.size(); d++) {
        CU_TRY(h, cudaSetDevice(h->slots[d].device));
        CU_TRY(h, cudaDeviceSynchronize());
        unsigned long long tmp[2];
        CU_TRY(h, cudaMemcpyFromSymbol(tmp, g_search_fused, sizeof tmp));
        out[0] += tmp[0];
        out[1] += tmp[1];
        if (reset) {
            unsigned long long z[2] = {0, 0};
            CU_TRY(h, cudaMemcpyToSymbol(g_search_fused, z, sizeof z));
        }
    }
    return 0;
}

int npswf_get_mf_calib(const npswf_handle *h, double *mfyref, double *mfint)
{
    if (!h || !mfyref || !mfint) return NPSWF_ERR_ARG;
    std::memcpy(mfyref, h->mfyref.data(), h->mfyref.size() * sizeof(double));
    std::memcpy(mfint, h->mfint.data(), h->mfint.size() * sizeof(double));
    return 0;
}

int npswf_get_spline(const npswf_handle *h, double *coef)
{
    if (!h || !coef) return NPSWF_ERR_ARG;
    std::memcpy(coef, h->spline.data(), h->spline.size() * sizeof(double));
    return 0;
}

const double *npswf_device_spline(const npswf_handle *h, int32_t dev_slot)
{
    if (!h || dev_slot < 0 || dev_slot >= (int)h->slots.size()) return nullptr;
    return h->slots[dev_slot].cal.spline;
}
const double *npswf_device_timeref(const npswf_handle *h, int32_t dev_slot)
{
    if (!h || dev_slot < 0 || dev_slot >= (int)h->slots.size()) return nullptr;
    return h->slots[dev_slot].cal.timeref;
}

int64_t npswf_flatten_event(const int32_t *wfnpulse, const double *wftime_padded, const double *wfampl_padded,
                            double *wftime_flat, double *wfampl_flat, int32_t *block_offset)
{
    int64_t off = 0;
    for (int i = 0; i < B; i++) {
        if (block_offset) block_offset[i] = (int32_t)off;  // T2:959
        const int n = wfnpulse[i];
        for (int p = 0; p < n && p < MAXP; p++) {
            if (wftime_flat) wftime_flat[off + p] = wftime_padded[(size_t)i * MAXP + p];
            if (wfampl_flat) wfampl_flat[off + p] = wfampl_padded[(size_t)i * MAXP + p];
        }
        off += (n < MAXP ? n : MAXP);  // T2:961
    }
    if (block_offset) block_offset[B] = (int32_t)off;  // T2:1022
    return off;
}

}  // extern "C"
