// band.cu -- halo exchange of the band decomposition behind the C ABI (include/pdegpu.h: pdegpu_band_*).
//
// One very large image is cut into column bands, one per GPU (BASELINE configs[4]); every T red-black sweeps
// a band needs the outermost H = 2T owned columns of its neighbours' unknowns in its halo columns. Here the exchange is
// two kernels on the context's stream and NO host synchronisation, NCCL call or stream event per step:
//
//   push  copies this band's outermost H owned columns of every unknown straight into the NEIGHBOUR'S mailbox (peer
//         stores over NVLink), then publishes the step number in the neighbour's memory (system-scope release);
//   pull  waits (device side, system-scope acquire) until the step's data has arrived in the own mailboxes, copies it
//         into the halo columns and acknowledges in the neighbour's memory, which is what allows the neighbour's NEXT
//         push to overwrite the mailbox.
//
// A mailbox block is one cudaMalloc owned by the library: [data from the left][data from the right][flags]. It is
// reached by the neighbours either through CUDA IPC handles (one process per GPU: pdegpu_band_export / _connect) or
// directly (several contexts in one process: pdegpu_band_connect_local). Waiting kernels never wait on something that
// comes later in their own stream: push(s) needs the neighbour's pull(s-1), pull(s) the neighbour's push(s), and every
// rank enqueues push(s) before pull(s) -- so two ranks cannot wait for each other. Spins are bounded (trap, not hang).
#include "pdegpu_internal.cuh"
#include <stdint.h>

struct pdegpu_band {
    pdegpu_ctx *ctx;
    int nrows, H, nunk, has_left, has_right;
    size_t side_floats;              // floats of one mailbox: nunk * H * nrows
    char *block;                     // own mailbox block (device memory)
    size_t block_bytes;
    char *peer[2];                   // neighbours' blocks as seen from this device (left, right); null where there is none
    int peer_ipc[2];                 // opened with cudaIpcOpenMemHandle (to be closed)
    unsigned step;
    unsigned long long bytes_sent;
};

namespace {

// flags at the end of a block
enum { F_DATA_FROM_LEFT = 0, F_DATA_FROM_RIGHT = 1, F_ACK_FROM_LEFT = 2, F_ACK_FROM_RIGHT = 3, F_DONE_CTAS = 4, F_COUNT = 16 };

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_ge_sys(const unsigned *p, unsigned want)
{
    unsigned long long spins = 0;
    while (ld_acquire_sys(p) < want) {
        __nanosleep(200);
        if (++spins > (1ull << 26)) __trap();                 // ~ 15 s: the neighbour is gone
    }
}

struct PushArgs {
    const float *src[2][2];          // [side][unknown]: first of the H columns to send
    float *dst[2];                   // [side]: mailbox in the neighbour's block
    unsigned *dst_flag[2];           // [side]: the neighbour's "data from ..." flag
    const unsigned *ack[2];          // [side]: own "ack from ..." flag
    unsigned *done;                  // own CTA counter
    int nunk, n4;                    // floats / 4 per unknown and side (H * nrows / 4)
    unsigned step;
};

// grid: (CTAs per side, 2 sides). Every CTA waits for the acknowledgement of the previous step, copies its share, and
// the last CTA of a side to finish publishes the step.
__global__ void __launch_bounds__(256)
band_push_kernel(const PushArgs a)
{
    const int side = blockIdx.y;
    if (!a.dst[side]) return;
    if (threadIdx.x == 0) wait_ge_sys(a.ack[side], a.step - 1);
    __syncthreads();
    for (int q = 0; q < a.nunk; q++) {
        const float4 *s = reinterpret_cast<const float4 *>(a.src[side][q]);
        float4 *d = reinterpret_cast<float4 *>(a.dst[side]) + (size_t)q * a.n4;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += gridDim.x * blockDim.x) d[i] = s[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(a.done + side, 1u);
        if (prev == gridDim.x - 1) {
            a.done[side] = 0;
            __threadfence_system();
            st_release_sys(a.dst_flag[side], a.step);
        }
    }
}

struct PullArgs {
    float *dst[2][2];                // [side][unknown]: first halo column
    const float *src[2];             // [side]: own mailbox
    const unsigned *flag[2];         // [side]: own "data from ..." flag
    unsigned *ack[2];                // [side]: the neighbour's "ack from ..." flag
    unsigned *done;
    int nunk, n4;
    unsigned step;
};

__global__ void __launch_bounds__(256)
band_pull_kernel(const PullArgs a)
{
    const int side = blockIdx.y;
    if (!a.src[side]) return;
    if (threadIdx.x == 0) wait_ge_sys(a.flag[side], a.step);
    __syncthreads();
    for (int q = 0; q < a.nunk; q++) {
        const float4 *s = reinterpret_cast<const float4 *>(a.src[side]) + (size_t)q * a.n4;
        float4 *d = reinterpret_cast<float4 *>(a.dst[side][q]);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += gridDim.x * blockDim.x) d[i] = __ldcv(s + i);   // (written by the peer: never from L1)
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(a.done + 2 + side, 1u);
        if (prev == gridDim.x - 1) {
            a.done[2 + side] = 0;
            __threadfence_system();
            st_release_sys(a.ack[side], a.step);
        }
    }
}

inline unsigned *flags_of(char *block, size_t side_floats) { return reinterpret_cast<unsigned *>(block + 2 * side_floats * sizeof(float)); }

}  // namespace

extern "C" int pdegpu_band_create(pdegpu_ctx *ctx, int nrows, int halo_cols, int nunk, int has_left, int has_right, pdegpu_band **out)
{
    if (!ctx || !out) return PDEGPU_ERR_ARG;
    if (nrows < 4 || (nrows & 3) || halo_cols < 1 || nunk < 1 || nunk > 2)
        return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_band_create: nrows must be a multiple of 4, 1 <= halo_cols, nunk in {1,2}");
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    pdegpu_band *b = (pdegpu_band *)calloc(1, sizeof(pdegpu_band));
    if (!b) return pdegpu_set_error(ctx, PDEGPU_ERR_NOMEM, "pdegpu_band_create: out of host memory");
    b->ctx = ctx; b->nrows = nrows; b->H = halo_cols; b->nunk = nunk; b->has_left = has_left != 0; b->has_right = has_right != 0;
    b->side_floats = (size_t)nunk * halo_cols * nrows;
    b->block_bytes = 2 * b->side_floats * sizeof(float) + F_COUNT * sizeof(unsigned);
    cudaError_t e = cudaMalloc((void **)&b->block, b->block_bytes);
    if (e != cudaSuccess) { free(b); return pdegpu_check_cuda(ctx, e, "cudaMalloc(band mailbox)"); }
    e = cudaMemset(b->block, 0, b->block_bytes);
    if (e != cudaSuccess) { cudaFree(b->block); free(b); return pdegpu_check_cuda(ctx, e, "cudaMemset(band mailbox)"); }
    *out = b;
    return PDEGPU_OK;
}

extern "C" int pdegpu_band_export(pdegpu_band *b, void *handle64)
{
    if (!b || !handle64) return PDEGPU_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    PDEGPU_CUDA_OK(b->ctx, cudaSetDevice(b->ctx->device));
    cudaIpcMemHandle_t h;
    PDEGPU_CUDA_OK(b->ctx, cudaIpcGetMemHandle(&h, b->block));
    memcpy(handle64, &h, sizeof h);
    return PDEGPU_OK;
}

extern "C" int pdegpu_band_connect(pdegpu_band *b, const void *left_handle64, const void *right_handle64)
{
    if (!b) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(b->ctx, cudaSetDevice(b->ctx->device));
    const void *hs[2] = {left_handle64, right_handle64};
    const int need[2] = {b->has_left, b->has_right};
    for (int s = 0; s < 2; s++) {
        if (!need[s]) continue;
        if (!hs[s]) return pdegpu_set_error(b->ctx, PDEGPU_ERR_ARG, "pdegpu_band_connect: missing handle of the %s neighbour", s ? "right" : "left");
        cudaIpcMemHandle_t h;
        memcpy(&h, hs[s], sizeof h);
        void *p = nullptr;
        PDEGPU_CUDA_OK(b->ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        b->peer[s] = (char *)p; b->peer_ipc[s] = 1;
    }
    return PDEGPU_OK;
}

extern "C" int pdegpu_band_connect_local(pdegpu_band *b, pdegpu_band *left, pdegpu_band *right)
{
    if (!b) return PDEGPU_ERR_ARG;
    PDEGPU_CUDA_OK(b->ctx, cudaSetDevice(b->ctx->device));
    pdegpu_band *nb[2] = {left, right};
    const int need[2] = {b->has_left, b->has_right};
    for (int s = 0; s < 2; s++) {
        if (!need[s]) continue;
        if (!nb[s] || nb[s]->side_floats != b->side_floats)
            return pdegpu_set_error(b->ctx, PDEGPU_ERR_ARG, "pdegpu_band_connect_local: %s neighbour missing or of another shape", s ? "right" : "left");
        if (nb[s]->ctx->device != b->ctx->device) {
            int can = 0;
            PDEGPU_CUDA_OK(b->ctx, cudaDeviceCanAccessPeer(&can, b->ctx->device, nb[s]->ctx->device));
            if (!can) return pdegpu_set_error(b->ctx, PDEGPU_ERR_UNSUPPORTED, "pdegpu_band_connect_local: no peer access from device %d to %d", b->ctx->device, nb[s]->ctx->device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(nb[s]->ctx->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return pdegpu_check_cuda(b->ctx, e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
        b->peer[s] = nb[s]->block; b->peer_ipc[s] = 0;
    }
    return PDEGPU_OK;
}

// unknowns[q]: this band's local array of unknown q, column-major with `nrows` rows, halo columns included;
// own0 / own1: first owned column / one past the last owned column (local indices)
extern "C" int pdegpu_band_exchange(pdegpu_band *b, float *const unknowns[], int own0, int own1)
{
    if (!b || !unknowns) return PDEGPU_ERR_ARG;
    pdegpu_ctx *ctx = b->ctx;
    if ((b->has_left && !b->peer[0]) || (b->has_right && !b->peer[1])) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_band_exchange: not connected");
    if (!b->has_left && !b->has_right) return PDEGPU_OK;
    const int H = b->H;
    if (own1 - own0 < H || (b->has_left && own0 < H)) return pdegpu_set_error(ctx, PDEGPU_ERR_SHAPE, "pdegpu_band_exchange: band narrower than the halo");
    for (int q = 0; q < b->nunk; q++)
        if (!unknowns[q] || ((uintptr_t)unknowns[q] & 15)) return pdegpu_set_error(ctx, PDEGPU_ERR_ARG, "pdegpu_band_exchange: unknowns must be 16-byte aligned");
    PDEGPU_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    b->step++;
    unsigned *my = flags_of(b->block, b->side_floats);
    const size_t col = (size_t)b->nrows;
    PushArgs pa;
    PullArgs ua;
    memset(&pa, 0, sizeof pa); memset(&ua, 0, sizeof ua);
    pa.nunk = ua.nunk = b->nunk; pa.n4 = ua.n4 = (int)((size_t)H * b->nrows / 4); pa.step = ua.step = b->step;
    pa.done = ua.done = my + F_DONE_CTAS;
    if (b->has_left) {
        unsigned *theirs = flags_of(b->peer[0], b->side_floats);
        for (int q = 0; q < b->nunk; q++) { pa.src[0][q] = unknowns[q] + (size_t)own0 * col; ua.dst[0][q] = unknowns[q] + (size_t)(own0 - H) * col; }
        pa.dst[0] = reinterpret_cast<float *>(b->peer[0]) + b->side_floats;       // I am its RIGHT neighbour
        pa.dst_flag[0] = theirs + F_DATA_FROM_RIGHT; pa.ack[0] = my + F_ACK_FROM_LEFT;
        ua.src[0] = reinterpret_cast<float *>(b->block); ua.flag[0] = my + F_DATA_FROM_LEFT; ua.ack[0] = theirs + F_ACK_FROM_RIGHT;
    }
    if (b->has_right) {
        unsigned *theirs = flags_of(b->peer[1], b->side_floats);
        for (int q = 0; q < b->nunk; q++) { pa.src[1][q] = unknowns[q] + (size_t)(own1 - H) * col; ua.dst[1][q] = unknowns[q] + (size_t)own1 * col; }
        pa.dst[1] = reinterpret_cast<float *>(b->peer[1]);                         // I am its LEFT neighbour
        pa.dst_flag[1] = theirs + F_DATA_FROM_LEFT; pa.ack[1] = my + F_ACK_FROM_RIGHT;
        ua.src[1] = reinterpret_cast<float *>(b->block) + b->side_floats; ua.flag[1] = my + F_DATA_FROM_RIGHT; ua.ack[1] = theirs + F_ACK_FROM_LEFT;
    }
    int ctas = (pa.n4 + 255) / 256;
    if (ctas > 32) ctas = 32;                                  // all CTAs of a launch must be resident at once (they wait)
    PDEGPU_PROF(ctx, "band_push_kernel", 0);
    band_push_kernel<<<dim3(ctas, 2), 256, 0, ctx->stream>>>(pa);
    PDEGPU_LAUNCH_CHECK(ctx, "band_push_kernel");
    PDEGPU_PROF(ctx, "band_pull_kernel", 0);
    band_pull_kernel<<<dim3(ctas, 2), 256, 0, ctx->stream>>>(ua);
    PDEGPU_LAUNCH_CHECK(ctx, "band_pull_kernel");
    b->bytes_sent += (unsigned long long)(b->has_left + b->has_right) * b->side_floats * sizeof(float);
    return PDEGPU_OK;
}

extern "C" unsigned long long pdegpu_band_bytes_sent(const pdegpu_band *b) { return b ? b->bytes_sent : 0ull; }

extern "C" void pdegpu_band_free(pdegpu_band *b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    for (int s = 0; s < 2; s++) if (b->peer[s] && b->peer_ipc[s]) cudaIpcCloseMemHandle(b->peer[s]);
    if (b->block) cudaFree(b->block);
    free(b);
}
