// link.cu -- the per-step row exchange of two neighbouring slabs over NVLink peer memory.
//
// The reference has no multi-device mode; its distributed users exchange ghost
// layers outside the library.  The slab decomposition of this library needs, every
// step, the rows of the points in the outermost owned cell layer (ghosts of the neighbour) and of
// the points that left the slab (migrants) on the neighbouring rank.  With NCCL that is two
// send/recv rounds (counts, then rows) with a host read-back between them; measured on 8 B200 it
// costs ~0.9 ms of latency per step, more than update! + the payload gather can hide.
//
// Here ONE kernel classifies the owned points (cell layer of the last dimension, full_grid.jl:93
// arithmetic), packs the rows of the leaving / boundary points and stores them STRAIGHT into the
// neighbour's receive buffer (a cudaIpc mapping of its memory, NVLink stores), a one-thread kernel
// behind it publishes (count, step number) in the neighbour's flag word, and the receiver's stream
// waits for that flag with a one-thread kernel.  No NCCL call, no count round trip; the host
// synchronises once per step (it needs the row counts as launch arguments).
//
// Memory of one rank (one cudaMalloc, exported with cudaIpcGetMemHandle):
//   flags[2 dirs][2 parities] (64 bytes each: step number, row count)
//   rows [2 dirs][2 parities][cap * width] floats     dir 0 = written by rank - 1, 1 = by rank + 1
// Two parities are enough: a rank sends step s + 2 only after it has received the neighbour's
// step s + 1 rows, which the neighbour sent after it had finished everything of step s.
#include <cstring>

#include "grid.cuh"

using namespace pnb;

struct LinkFlag {
    unsigned long long seq;
    long long n;
    long long pad_[6];
};

struct pnb_slab_link {
    int64_t cap;
    int width;                     // floats of a row
    int stride;                    // floats between rows in the receive buffers: width rounded up to 4
    unsigned char *area;           // my receive area
    size_t area_bytes;
    unsigned char *peer[2];        // mapped areas of rank - 1 (0) and rank + 1 (1), or nullptr
    int32_t *counts;               // device: n_down, n_up, n_leave of the current send
    long long *h_out;              // mapped pinned: what pnb_slab_link_recv reports
    long long *d_out;
    uint64_t seq_sent, seq_recv;
    bool peer_local[2];            // connected with pnb_slab_link_connect_local (same process)
    long long timeout_ns;          // how long a receive waits for a neighbour (default 120 s)
};

static size_t link_flags_bytes() { return 4 * sizeof(LinkFlag); }
static size_t link_buf_bytes(const pnb_slab_link *l) { return ((size_t)l->cap * l->stride * sizeof(float) + 255) / 256 * 256; }
static LinkFlag *link_flag(unsigned char *area, int dir, int par) { return reinterpret_cast<LinkFlag *>(area) + dir * 2 + par; }
static float *link_rows(const pnb_slab_link *l, unsigned char *area, int dir, int par)
{
    return reinterpret_cast<float *>(area + link_flags_bytes() + (size_t)(dir * 2 + par) * link_buf_bytes(l));
}

extern "C" void pnb_slab_link_destroy(pnb_slab_link *l)
{
    if (!l) return;
    cudaDeviceSynchronize();
    for (int d = 0; d < 2; d++)
        if (l->peer[d] && !l->peer_local[d]) cudaIpcCloseMemHandle(l->peer[d]);
    cudaFree(l->area);
    cudaFree(l->counts);
    if (l->h_out) cudaFreeHost(l->h_out);
    cudaGetLastError();
    delete l;
}

extern "C" pnb_status pnb_slab_link_create(int64_t cap_rows, int width, pnb_slab_link **out)
{
    if (!out) { set_error("out is NULL"); return PNB_ERR_ARG; }
    *out = nullptr;
    if (cap_rows <= 0 || width < 1 || width > 32) { set_error("link: cap_rows > 0 and 1 <= width <= 32"); return PNB_ERR_ARG; }
    pnb_slab_link *l = new pnb_slab_link();
    memset(l, 0, sizeof(*l));
    l->cap = cap_rows;
    l->width = width;
    l->stride = (width + 3) & ~3;
    l->timeout_ns = 120LL * 1000 * 1000 * 1000;
    l->area_bytes = link_flags_bytes() + 4 * link_buf_bytes(l);
    cudaError_t e;
    if ((e = cudaMalloc(&l->area, l->area_bytes)) != cudaSuccess ||
        (e = cudaMemset(l->area, 0, link_flags_bytes())) != cudaSuccess ||
        (e = cudaMalloc(&l->counts, 4 * sizeof(int32_t))) != cudaSuccess ||
        (e = cudaHostAlloc(&l->h_out, 8 * sizeof(long long), cudaHostAllocMapped)) != cudaSuccess ||
        (e = cudaHostGetDevicePointer(&l->d_out, l->h_out, 0)) != cudaSuccess) {
        pnb_status st = cuda_fail(e, "pnb_slab_link_create");
        pnb_slab_link_destroy(l);
        return st;
    }
    PNB_CUDA(cudaDeviceSynchronize());
    *out = l;
    return PNB_OK;
}

// how long pnb_slab_link_recv lets the GPU wait for a neighbour before it gives up with
// PNB_ERR_STATE instead of hanging (default 120 s)
extern "C" pnb_status pnb_slab_link_set_timeout(pnb_slab_link *l, double seconds)
{
    if (!l || !(seconds > 0.0)) { set_error("link: positive timeout expected"); return PNB_ERR_ARG; }
    l->timeout_ns = (long long)(seconds * 1e9);
    return PNB_OK;
}

// handle_out: 64 bytes (cudaIpcMemHandle_t) a neighbouring rank passes to pnb_slab_link_connect
extern "C" pnb_status pnb_slab_link_export(pnb_slab_link *l, void *handle_out)
{
    if (!l || !handle_out) { set_error("NULL argument"); return PNB_ERR_ARG; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    PNB_CUDA(cudaIpcGetMemHandle(&h, l->area));
    memcpy(handle_out, &h, sizeof(h));
    return PNB_OK;
}

// handles of rank - 1 (down) and rank + 1 (up); NULL = no neighbour on that side.  The
// neighbours must have been created with the same cap_rows and width.
extern "C" pnb_status pnb_slab_link_connect(pnb_slab_link *l, const void *handle_down, const void *handle_up)
{
    if (!l) { set_error("NULL argument"); return PNB_ERR_ARG; }
    const void *hs[2] = {handle_down, handle_up};
    for (int d = 0; d < 2; d++) {
        if (!hs[d] || l->peer[d]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs[d], sizeof(h));
        void *p = nullptr;
        PNB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        l->peer[d] = static_cast<unsigned char *>(p);
    }
    return PNB_OK;
}

// Two links of the SAME process (two slabs on one device, or on two devices of one process with
// peer access enabled): plain pointers instead of cudaIpc handles.  Used by the single-GPU tests
// and profiles of the exchange kernels; the protocol is the same.
extern "C" pnb_status pnb_slab_link_connect_local(pnb_slab_link *l, pnb_slab_link *down, pnb_slab_link *up)
{
    if (!l) { set_error("NULL argument"); return PNB_ERR_ARG; }
    pnb_slab_link *o[2] = {down, up};
    for (int d = 0; d < 2; d++) {
        if (!o[d] || l->peer[d]) continue;
        if (o[d]->cap != l->cap || o[d]->width != l->width) {
            set_error("link: neighbours must be created with the same capacity and row width");
            return PNB_ERR_ARG;
        }
        l->peer[d] = o[d]->area;
        l->peer_local[d] = true;
    }
    return PNB_OK;
}

namespace pnb {

constexpr int kLinkMaxW = 16;      // rows up to this width are staged per warp for coalesced stores

// One pass over the owned points: class of every point (as k_slab_classify), the rows of the
// down / up classes go straight to the neighbours' receive buffers, the leavers' indices to
// leave_idx.  counts = {n_down, n_up, n_leave} (zeroed before the launch).
__global__ void __launch_bounds__(256)
k_link_classify_send(pnb_slab_arrays A, int WS, int64_t n, int nd, float pmin, float cs, float inv_cs, long long z_lo,
                     long long z_hi, float *__restrict__ dst_down, float *__restrict__ dst_up, int64_t cap,
                     int32_t *__restrict__ leave_idx, int64_t leave_cap, int32_t *__restrict__ counts)
{
    __shared__ __align__(16) float stage[8][32 * kLinkMaxW];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool up = false, down = false, leave = false;
    if (i < n) {
        const float z = __ldg(A.ptr[0] + i * nd + (nd - 1));
        // cheap test first (multiplication, a few ulp off): a point clearly strictly inside the
        // owned layers (z_lo < layer < z_hi) is neither sent nor leaving; only the others pay for
        // the exact division of full_grid.jl:93
        const float fa = __fsub_rn(z, pmin) * inv_cs;
        const float tol = 1e-3f + 1e-5f * fabsf(fa);
        if (!(fa > (float)z_lo + tol && fa < (float)(z_hi - 1) - tol)) {
            const float f = floorf(__fdiv_rn(__fsub_rn(z, pmin), cs));
            if (fabsf(f) < 4.0e18f) {          // NaN / huge values stay: update! reports them
                const long long cz = (long long)f + 1;
                up = dst_up != nullptr && cz >= z_hi;
                down = dst_down != nullptr && cz <= z_lo;
                leave = (cz < z_lo && dst_down != nullptr) || (cz > z_hi && dst_up != nullptr);
            }
        }
    }
    if (!__any_sync(0xffffffffu, up || down || leave)) return;
#pragma unroll
    for (int dir = 0; dir < 2; dir++) {
        const bool flag = dir ? up : down;
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        if (m == 0u) continue;
        float *dst = dir ? dst_up : dst_down;
        const int cnt = __popc(m);
        int base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(counts + dir, cnt);
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        const int rank = __popc(m & ((1u << lane) - 1u));
        long long n_ok = (long long)cap - base;                 // rows of this warp that fit
        n_ok = n_ok < 0 ? 0 : (n_ok > cnt ? cnt : n_ok);
        if (WS <= kLinkMaxW) {
            if (flag) {
                int col = 0;
#pragma unroll
                for (int a = 0; a < 8; a++) {           // static indices: A stays in the parameter bank
                    if (a < A.n_arrays) {
                        const int w = A.width[a];
                        for (int k = 0; k < w; k++) stage[warp][rank * WS + col + k] = __ldg(A.ptr[a] + i * w + k);
                        col += w;
                    }
                }
                for (; col < WS; col++) stage[warp][rank * WS + col] = 0.0f;
            }
            __syncwarp();
            // rows are padded to a multiple of 16 bytes: aligned 16-byte NVLink stores
            float4 *o = reinterpret_cast<float4 *>(dst + (int64_t)base * WS);
            const float4 *st4 = reinterpret_cast<const float4 *>(stage[warp]);
            for (int j = lane; j < (int)n_ok * (WS >> 2); j += 32) o[j] = st4[j];
            __syncwarp();
        } else if (flag && rank < n_ok) {
            float *o = dst + ((int64_t)base + rank) * WS;
            int col = 0;
#pragma unroll
            for (int a = 0; a < 8; a++) {
                if (a < A.n_arrays) {
                    const int w = A.width[a];
                    for (int k = 0; k < w; k++) o[col + k] = __ldg(A.ptr[a] + i * w + k);
                    col += w;
                }
            }
        }
    }
    {
        const unsigned m = __ballot_sync(0xffffffffu, leave);
        if (m) {
            int base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(counts + 2, __popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (leave) {
                const int64_t pos = (int64_t)base + __popc(m & ((1u << lane) - 1u));
                if (pos < leave_cap) leave_idx[pos] = (int32_t)i;
            }
        }
    }
}

// behind k_link_classify_send on the same stream: all rows are in the neighbours' memory (kernel
// boundary), now the counts and, last, the step number
__global__ void k_link_publish(const int32_t *__restrict__ counts, LinkFlag *flag_down, LinkFlag *flag_up,
                               unsigned long long seq)
{
    __threadfence_system();
    LinkFlag *f[2] = {flag_down, flag_up};
    for (int d = 0; d < 2; d++) {
        if (!f[d]) continue;
        *(volatile long long *)&f[d]->n = counts[d];
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&f[d]->seq), "l"(seq) : "memory");
    }
}

// waits until both neighbours have published step `seq`; out = {n_from_down, n_from_up, n_down,
// n_up, n_leave, status}: status 1 = a neighbour did not answer within the time limit
__global__ void k_link_wait(const LinkFlag *flag_down, const LinkFlag *flag_up, unsigned long long seq,
                            const int32_t *__restrict__ counts, long long *out, long long timeout_ns)
{
    const LinkFlag *f[2] = {flag_down, flag_up};
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    long long status = 0;
    for (int d = 0; d < 2; d++) {
        long long n = 0;
        if (f[d]) {
            for (;;) {
                unsigned long long s;
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(s) : "l"(&f[d]->seq) : "memory");
                if (s == seq) break;
                unsigned long long t;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                if ((long long)(t - t0) > timeout_ns) { status = 1; break; }
                __nanosleep(200);
            }
            n = *(volatile const long long *)&f[d]->n;
        }
        out[d] = n;
    }
    out[2] = counts[0];
    out[3] = counts[1];
    out[4] = counts[2];
    out[5] = status;
    __threadfence_system();
}

}  // namespace pnb

// Classify the n owned points of `arrays` (array 0 = coordinates) and send the rows of step `seq`
// (1, 2, 3, ... -- the same number on all ranks) to the neighbours.  leave_idx: >= 2 * cap int32.
// Stream-ordered, nothing is synchronised.
extern "C" pnb_status pnb_slab_link_send(pnb_slab_link *l, const pnb_slab_arrays *arrays, int64_t n, int ndims,
                                         float padded_min_z, float cell_size_z, int64_t z_lo, int64_t z_hi,
                                         int32_t *leave_idx, uint64_t seq, void *stream)
{
    if (!l || !arrays || !leave_idx) { set_error("NULL argument"); return PNB_ERR_ARG; }
    if (arrays->n_arrays < 1 || arrays->n_arrays > 8 || arrays->width[0] != ndims) {
        set_error("arrays[0] must be the coordinates (width = NDIMS), at most 8 arrays");
        return PNB_ERR_ARG;
    }
    int W = 0;
    for (int a = 0; a < arrays->n_arrays; a++) W += arrays->width[a];
    if (W != l->width) { set_error("link: row width %d, created for %d", W, l->width); return PNB_ERR_ARG; }
    if (n < 0 || n > 0x7ffffff0LL) { set_error("link: 0 <= n < 2^31 points expected"); return PNB_ERR_ARG; }
    if (seq != l->seq_sent + 1) { set_error("link: step %llu sent after step %llu", (unsigned long long)seq, (unsigned long long)l->seq_sent); return PNB_ERR_STATE; }
    cudaStream_t s = (cudaStream_t)stream;
    const int par = (int)(seq & 1);
    // my rows for rank - 1 land in ITS "written by rank + 1" buffers (dir 1), and vice versa
    float *dst_down = l->peer[0] ? link_rows(l, l->peer[0], 1, par) : nullptr;
    float *dst_up = l->peer[1] ? link_rows(l, l->peer[1], 0, par) : nullptr;
    PNB_CUDA(cudaMemsetAsync(l->counts, 0, 4 * sizeof(int32_t), s));
    if (n > 0) {
        k_link_classify_send<<<(unsigned)div_up(n, 256), 256, 0, s>>>(
            *arrays, l->stride, n, ndims, padded_min_z, cell_size_z, 1.0f / cell_size_z, (long long)z_lo, (long long)z_hi, dst_down,
            dst_up, l->cap, leave_idx, 2 * l->cap, l->counts);
        PNB_LAUNCHED();
    }
    k_link_publish<<<1, 1, 0, s>>>(l->counts, l->peer[0] ? link_flag(l->peer[0], 1, par) : nullptr,
                                   l->peer[1] ? link_flag(l->peer[1], 0, par) : nullptr,
                                   (unsigned long long)seq);
    PNB_LAUNCHED();
    l->seq_sent = seq;
    return PNB_OK;
}

// Wait for the rows of step `seq` from both neighbours.  SYNCHRONISES the stream (the host needs
// the counts).  rows_down / rows_up: device pointers into this rank's receive area (valid until
// step seq + 2 is received); counts = {n_from_down, n_from_up, n_sent_down, n_sent_up, n_leave}.
extern "C" int pnb_slab_link_row_stride(const pnb_slab_link *l) { return l ? l->stride : 0; }

extern "C" pnb_status pnb_slab_link_recv(pnb_slab_link *l, uint64_t seq, const float **rows_down,
                                         const float **rows_up, int64_t *counts, void *stream)
{
    if (!l || !rows_down || !rows_up || !counts) { set_error("NULL argument"); return PNB_ERR_ARG; }
    if (seq != l->seq_sent) { set_error("link: receive of step %llu before its send", (unsigned long long)seq); return PNB_ERR_STATE; }
    cudaStream_t s = (cudaStream_t)stream;
    const int par = (int)(seq & 1);
    k_link_wait<<<1, 1, 0, s>>>(l->peer[0] ? link_flag(l->area, 0, par) : nullptr,
                                l->peer[1] ? link_flag(l->area, 1, par) : nullptr, (unsigned long long)seq,
                                l->counts, l->d_out, l->timeout_ns);
    PNB_LAUNCHED();
    PNB_CUDA(cudaStreamSynchronize(s));
    l->seq_recv = seq;
    volatile long long *h = l->h_out;
    for (int k = 0; k < 5; k++) counts[k] = h[k];
    if (h[5] != 0) {
        set_error("link: a neighbouring rank did not send step %llu within %.0f s (pnb_slab_link_set_timeout)",
                  (unsigned long long)seq, 1e-9 * (double)l->timeout_ns);
        return PNB_ERR_STATE;
    }
    if (counts[0] > l->cap || counts[1] > l->cap || counts[2] > l->cap || counts[3] > l->cap) {
        set_error("link: %lld / %lld rows received, %lld / %lld sent, capacity %lld rows: create the link "
                  "with a larger capacity", (long long)counts[0], (long long)counts[1], (long long)counts[2],
                  (long long)counts[3], (long long)l->cap);
        return PNB_ERR_LIST_FULL;
    }
    *rows_down = link_rows(l, l->area, 0, par);
    *rows_up = link_rows(l, l->area, 1, par);
    return PNB_OK;
}
