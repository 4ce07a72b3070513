// svgf_band.cu — include/svgf_band.h: one frame in horizontal bands, one band per GPU.  The frame of a band is a resumable
// state machine (frame_begin / frame_run / frame_exchange) that stops at every exchange; three transports perform it:
// NCCL send/recv, peer-memory pulls between processes (CUDA IPC mappings + flag words + one pull kernel per exchange), and
// peer copies between the bands of an in-process group.
//
// What runs where, per frame (N = 5 levels; main = the caller's stream, side = the driver's exchange stream):
//   main: [wait STATE(t-1)]  temporal + variance (whole local image)  level 0 (whole local image)
//   main: level 1 (band +- 8 rows)   level 2: both boundary strips -> [ev]   level 2: interior row blocks
//   side:                                                               `-> exchange 16 rows of level 2's output -> [ev]
//   main:                       [wait]  level 3: both boundary strips -> [ev]   level 3: interior row blocks
//   side:                                                                 `-> exchange 32 rows of level 3's output + STATE(t) -> [ev]
//   main:       level 4: the row blocks that read no apron row   [wait]  level 4: the others -> filter[0]
//   frame t+1, main: [wait STATE(t)]
// (NCCL and the in-process group send STATE(t) with the last halo exchange as drawn; the peer-memory transport pulls it with
// a kernel of its own right after level 0.)
// Levels 0..2 do not exchange: their halos (2 + 4 + 8 rows), the variance pass's 7x7 window and the motion vectors' reach are
// covered by computing those stages up to 17 + max_motion rows into the 32-row apron.  Only lattice planes travel for
// the per-level exchanges (the level that follows reads them by TMA); whole padded rows, so one contiguous block per plane.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): libsvgf_b200.so itself has no NCCL dependency, and inside a PyTorch
// process the call lands in the library PyTorch has already loaded.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <new>

#include "../../include/svgf_band.h"
#include "svgf_ctx.h"

namespace {

// ---- the slice of the NCCL API used here (nccl.h: stable since 2.7) ----------------------------------------------------
struct NcclUniqueId { char internal[SVGF_BAND_UNIQUE_ID_BYTES]; };
typedef struct ncclComm *NcclComm;
constexpr int kNcclUint8 = 1;   // ncclDataType_t::ncclUint8
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    bool ok = false;
};

const NcclApi &nccl() {
    static NcclApi api = [] {
        NcclApi a;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) return a;
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.lib, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.lib, "ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
        a.Send = reinterpret_cast<decltype(a.Send)>(dlsym(a.lib, "ncclSend"));
        a.Recv = reinterpret_cast<decltype(a.Recv)>(dlsym(a.lib, "ncclRecv"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(a.lib, "ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(a.lib, "ncclGroupEnd"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Send && a.Recv && a.GroupStart && a.GroupEnd;
        return a;
    }();
    return api;
}

struct Plane {          // rows of a local image as one contiguous run per row range
    char *base;         // address of local row 0
    size_t row_bytes;
};
struct PlaneSet { const Plane *planes; int n; int rows; };

// ---- the peer-memory transport (svgf_band_create_ipc): flags and pulls over NVLink, no NCCL ---------------------------------
// Every rank exports its lattice colour planes, a staging block for the state rows and a few flag words (cudaIpc handles);
// a neighbour maps them.  An exchange, on the consumer's side stream: publish READY and wait for the producers' READY flags
// (one thread), copy their rows over NVLink into the local aprons (one grid), tell them the rows have been read (PULLED
// flags, so that a producer never overwrites rows a slower neighbour has not fetched yet).  Flags are monotonically
// increasing sequence numbers.
struct PullJob { const char *src; char *dst; unsigned long long bytes; };
struct FlagRef { unsigned *p; unsigned v; };
struct PullArgs {
    PullJob job[12];
    FlagRef signal[4];
    unsigned *counter;
    int njobs, nsignal;
};
struct FlagArgs { FlagRef publish[2], wait[4]; int npublish, nwait; };

constexpr long long kSpinLimitClocks = 20000000000LL;      // ~10 s: a neighbour that never arrives becomes a CUDA error, not a hang

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ void spin_until_reached(const unsigned *p, unsigned v) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(p) - v) < 0) {
        __nanosleep(100);
        if (clock64() - t0 > kSpinLimitClocks) __trap();
    }
}

// The copy of one exchange.  It is launched behind band_flag_kernel, which has done the waiting with ONE thread: a grid that
// spins occupies the SM slots the level kernels on the main stream need (64 waiting CTAs cost the interior of a level a
// fifth of the GPU for as long as a neighbour lags - measured at N = 4: 1.10 ms against 0.97 ms).
__global__ void __launch_bounds__(256) band_pull_kernel(PullArgs a) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (int j = 0; j < a.njobs; j++) {
        const PullJob q = a.job[j];
        if (((reinterpret_cast<size_t>(q.src) | reinterpret_cast<size_t>(q.dst) | q.bytes) & 15) == 0) {
            const uint4 *src = reinterpret_cast<const uint4 *>(q.src);
            uint4 *dst = reinterpret_cast<uint4 *>(q.dst);
            const size_t n = q.bytes >> 4;
            size_t k = tid;
            for (; k + 3 * nth < n; k += 4 * nth) {        // four loads in flight per thread: the round trip is an NVLink hop
                const uint4 v0 = __ldcv(src + k), v1 = __ldcv(src + k + nth), v2 = __ldcv(src + k + 2 * nth), v3 = __ldcv(src + k + 3 * nth);
                dst[k] = v0; dst[k + nth] = v1; dst[k + 2 * nth] = v2; dst[k + 3 * nth] = v3;
            }
            for (; k < n; k += nth) dst[k] = __ldcv(src + k);
        } else {                                            // history lengths of a width that is not a multiple of 16
            for (size_t k = tid; k < q.bytes; k += nth) q.dst[k] = __ldcv(q.src + k);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.counter, 1u) == gridDim.x - 1) {    // last CTA: every row has been read
            *a.counter = 0;
            __threadfence_system();
            for (int i = 0; i < a.nsignal; i++) st_release_sys(a.signal[i].p, a.signal[i].v);
        }
    }
}

// One thread: publish this band's READY numbers (the launch is ordered behind the kernels that wrote the rows), then wait for
// the neighbours' numbers (READY before a pull, PULLED at the start of a frame).
__global__ void band_flag_kernel(FlagArgs a) {
    if (a.npublish) __threadfence_system();
    for (int i = 0; i < a.npublish; i++) st_release_sys(a.publish[i].p, a.publish[i].v);
    for (int i = 0; i < a.nwait; i++) spin_until_reached(a.wait[i].p, a.wait[i].v);
}

struct IpcBlob {                       // what svgf_band_ipc_export hands to the neighbours (SVGF_BAND_IPC_BYTES)
    unsigned magic;
    int width, storage, band_lo, band_hi, local_rows, pitch_pairs;
    cudaIpcMemHandle_t planes[6], staging, flags;
};
constexpr unsigned kIpcMagic = 0x53424950u;   // "SBIP"
static_assert(sizeof(IpcBlob) <= SVGF_BAND_IPC_BYTES, "IpcBlob must fit the public blob size");

struct IpcPeer {
    bool open = false;
    char *planes[6] = {};              // lattice colour planes sc[0].{c0,c1,lz}, sc[1].{c0,c1,lz}
    char *staging = nullptr;
    unsigned *flags = nullptr;
    int band_lo = 0, band_hi = 0;
};
// flag words (each rank owns one array; 32 words apart so that two never share a line)
enum { kFlagReadyHalo = 0, kFlagReadyState = 1, kFlagPulledHalo = 2 /* +dir */, kFlagPulledState = 4 /* +dir */, kFlagWords = 6, kFlagStride = 32 };

}  // namespace

struct svgf_band {
    svgf_ctx *ctx = nullptr;
    int device = 0, rank = 0, world = 1, W = 0, H = 0;
    int y0 = 0, y1 = 0, ly0 = 0, ly1 = 0;
    NcclComm comm = nullptr;
    cudaStream_t side = nullptr;
    cudaStream_t side_state = nullptr; // peer-memory transport: the state rows travel on a stream of their own
    cudaEvent_t ev_l0 = nullptr, ev_boundary[2] = {nullptr, nullptr}, ev_halo[2] = {nullptr, nullptr}, ev_state = nullptr, ev_pulled = nullptr;
    bool state_pending = false;
    bool dry_run = false;          // SVGF_FLAG_BAND_NO_EXCHANGE of the current call
    int last_err = 0;
    // in-process group (svgf_band_create_group): the neighbours' bands; halos are copied from their planes
    bool in_group = false;
    svgf_band *up = nullptr, *down = nullptr;
    // peer-memory transport (svgf_band_create_ipc)
    bool ipc = false, ipc_connected = false;
    IpcPeer peer[2];                   // 0 = up (rank - 1), 1 = down (rank + 1)
    char *staging = nullptr;           // [side 0 = my top band rows, 1 = bottom][colour, moments, history] of the last frame
    size_t staging_side_bytes = 0;
    unsigned *flags = nullptr, *pull_counter = nullptr, *pull_counter_state = nullptr;
    unsigned ticket = 0, halo_seq = 0; // frames begun / halo exchanges posted (the same on every rank)
    unsigned state_ticket = 0;         // ticket of the last frame whose state rows were published (frames without exchange skip it)
    uint64_t extra_launches = 0;       // flag and pull kernels (svgf_band_launch_count)

    // ---- the frame in flight: svgf_band_frame / svgf_band_group_frame advance it from exchange to exchange ----
    struct Frame {
        svgf_params params;
        svgf_frame_buffers bufs;
        cudaStream_t s = nullptr;
        int N = 0, slot = 0, src = 0, cur_level = 1, n_halo = 0, n_steps = 0, pc = 0, last_exchange = -1;
        svgf_band_step steps[32];
        Plane state[3];                // colour history, moments, history lengths of this frame (the next frame's previous state)
        // the exchange the frame is stopped at: up to two plane sets (lattice planes of a level, state planes)
        Plane lattice[3];
        PlaneSet sets[2];
        int n_sets = 0;
        cudaEvent_t ready = nullptr;   // recorded on the main stream: the rows this exchange sends are final
        cudaEvent_t done = nullptr;    // to record on the side stream after the exchange (null: none)
        bool carries_state = false;
    } f;

    int band_lo() const { return y0 - ly0; }     // local row of the first owned row
    int band_hi() const { return y1 - ly0; }     // local row one past the last owned row
    int local_rows() const { return ly1 - ly0; }
};

namespace {

svgf_status band_cuda(svgf_band *b, cudaError_t e) {
    if (e == cudaSuccess) return SVGF_OK;
    b->last_err = (int)e;
    return SVGF_CUDA_ERROR;
}
svgf_status band_nccl(svgf_band *b, int r) {
    if (r == 0) return SVGF_OK;
    b->last_err = 10000 + r;    // ncclResult_t, offset so that it cannot be mistaken for a cudaError_t
    return SVGF_CUDA_ERROR;
}
#define BAND_TRY(x)                       \
    do {                                  \
        const svgf_status st_ = (x);      \
        if (st_ != SVGF_OK) return st_;   \
    } while (0)

struct DeviceScope {
    int prev = -1, cur = -1;
    cudaError_t enter(int device) {
        cudaGetDevice(&prev);
        cur = device;
        return prev != device ? cudaSetDevice(device) : cudaSuccess;
    }
    ~DeviceScope() { if (prev >= 0 && prev != cur) cudaSetDevice(prev); }
};

// Refresh the apron rows on each side of the band with the neighbours' band rows, for every plane of up to two plane sets
// (each with its own row count): ONE grouped send/recv on the side stream.
svgf_status exchange(svgf_band *b, const PlaneSet *sets, int n_sets) {
    if (b->dry_run) return SVGF_OK;
    const NcclApi &n = nccl();
    const int lo = b->band_lo(), hi = b->band_hi();
    BAND_TRY(band_nccl(b, n.GroupStart()));
    for (int q = 0; q < n_sets; q++)
        for (int k = 0; k < sets[q].n; k++) {
            const Plane &p = sets[q].planes[k];
            const int rows = sets[q].rows;
            const size_t bytes = (size_t)rows * p.row_bytes;
            if (b->rank > 0) {                // my top band rows -> upper neighbour's bottom apron; its bottom band rows -> my top apron
                BAND_TRY(band_nccl(b, n.Send(p.base + (size_t)lo * p.row_bytes, bytes, kNcclUint8, b->rank - 1, b->comm, b->side)));
                BAND_TRY(band_nccl(b, n.Recv(p.base + (size_t)(lo - rows) * p.row_bytes, bytes, kNcclUint8, b->rank - 1, b->comm, b->side)));
            }
            if (b->rank + 1 < b->world) {
                BAND_TRY(band_nccl(b, n.Send(p.base + (size_t)(hi - rows) * p.row_bytes, bytes, kNcclUint8, b->rank + 1, b->comm, b->side)));
                BAND_TRY(band_nccl(b, n.Recv(p.base + (size_t)hi * p.row_bytes, bytes, kNcclUint8, b->rank + 1, b->comm, b->side)));
            }
        }
    return band_nccl(b, n.GroupEnd());
}

// The same refresh inside an in-process group: the neighbours' band rows are COPIED from their planes (same plane order,
// same row pitch) into this band's aprons on this band's side stream.  The caller has made the side stream wait for the
// neighbours' `ready` events.  cudaMemcpyPeerAsync: the bands may live on different devices or all on one.
svgf_status exchange_copy(svgf_band *b) {
    if (b->dry_run) return SVGF_OK;
    const int lo = b->band_lo(), hi = b->band_hi();
    for (int q = 0; q < b->f.n_sets; q++)
        for (int k = 0; k < b->f.sets[q].n; k++) {
            const Plane &p = b->f.sets[q].planes[k];
            const int rows = b->f.sets[q].rows;
            const size_t bytes = (size_t)rows * p.row_bytes;
            if (b->up) {
                const Plane &o = b->up->f.sets[q].planes[k];
                BAND_TRY(band_cuda(b, cudaMemcpyPeerAsync(p.base + (size_t)(lo - rows) * p.row_bytes, b->device,
                                                          o.base + (size_t)(b->up->band_hi() - rows) * o.row_bytes, b->up->device, bytes, b->side)));
            }
            if (b->down) {
                const Plane &o = b->down->f.sets[q].planes[k];
                BAND_TRY(band_cuda(b, cudaMemcpyPeerAsync(p.base + (size_t)hi * p.row_bytes, b->device,
                                                          o.base + (size_t)b->down->band_lo() * o.row_bytes, b->down->device, bytes, b->side)));
            }
        }
    return SVGF_OK;
}

svgf_status launch_flags(svgf_band *b, const FlagArgs &a, cudaStream_t s) {
    band_flag_kernel<<<1, 1, 0, s>>>(a);
    b->extra_launches++;
    return band_cuda(b, cudaGetLastError());
}

// Peer-memory transport, frame begin: the rows this frame is about to overwrite - boundary rows of the lattice planes, the
// state staging block - must have been fetched by the neighbours (they were, long ago, unless a rank lags).
svgf_status ipc_wait_pulled(svgf_band *b, cudaStream_t s) {
    FlagArgs a{};
    for (int d = 0; d < 2; d++) {
        if (!b->peer[d].open) continue;
        a.wait[a.nwait++] = FlagRef{b->flags + (kFlagPulledHalo + d) * kFlagStride, b->halo_seq};
        a.wait[a.nwait++] = FlagRef{b->flags + (kFlagPulledState + d) * kFlagStride, b->state_ticket};
    }
    return launch_flags(b, a, s);
}

// Peer-memory transport, after level 0 (side stream): this frame's state rows next to each neighbour -> staging, READY_STATE.
svgf_status ipc_publish_state(svgf_band *b) {
    const svgf_band::Frame &f = b->f;
    const int lo = b->band_lo(), hi = b->band_hi();
    for (int d = 0; d < 2; d++) {
        if (!b->peer[d].open) continue;
        char *dst = b->staging + (size_t)d * b->staging_side_bytes;
        const int row0 = d == 0 ? lo : hi - SVGF_BAND_APRON;
        for (int k = 0; k < 3; k++) {
            const size_t bytes = (size_t)SVGF_BAND_APRON * f.state[k].row_bytes;
            BAND_TRY(band_cuda(b, cudaMemcpyAsync(dst, f.state[k].base + (size_t)row0 * f.state[k].row_bytes, bytes, cudaMemcpyDeviceToDevice, b->side_state)));
            dst += bytes;
        }
    }
    return SVGF_OK;          // READY_STATE is published by the flag kernel of the pull that follows on the same stream
}

// Peer-memory transport, one exchange (side stream, after the event that says this band's rows are final): a one-thread
// kernel publishes READY for the neighbours and waits for theirs, then one grid pulls their rows into the aprons and
// acknowledges.
svgf_status exchange_ipc(svgf_band *b, bool has_halo, bool with_state) {
    if (b->dry_run) return SVGF_OK;
    svgf_band::Frame &f = b->f;
    cudaStream_t stream = with_state ? b->side_state : b->side;     // a call moves the halos or the state, never both
    const int lo = b->band_lo(), hi = b->band_hi();
    PullArgs a{};
    FlagArgs w{};
    a.counter = with_state ? b->pull_counter_state : b->pull_counter;
    if (has_halo) {
        b->halo_seq++;
        w.publish[w.npublish++] = FlagRef{b->flags + kFlagReadyHalo * kFlagStride, b->halo_seq};
    }
    if (with_state) w.publish[w.npublish++] = FlagRef{b->flags + kFlagReadyState * kFlagStride, b->ticket};
    unsigned long long total = 0;
    const size_t pad = has_halo ? (size_t)svgf::kLatPadY * f.lattice[0].row_bytes : 0;
    const int set_base = (1 - f.src) * 3;                  // the lattice colour set this level wrote, here and on the neighbours
    for (int d = 0; d < 2; d++) {
        const IpcPeer &p = b->peer[d];
        if (!p.open) continue;
        const int mine_at_peer = 1 - d;                    // the upper neighbour knows this band as its `down`
        if (has_halo) {
            const int rows = f.sets[0].rows;
            w.wait[w.nwait++] = FlagRef{p.flags + kFlagReadyHalo * kFlagStride, b->halo_seq};
            a.signal[a.nsignal++] = FlagRef{p.flags + (kFlagPulledHalo + mine_at_peer) * kFlagStride, b->halo_seq};
            for (int k = 0; k < 3; k++) {
                const size_t rb = f.lattice[k].row_bytes, bytes = (size_t)rows * rb;
                const char *src = p.planes[set_base + k] + pad + (size_t)(d == 0 ? p.band_hi - rows : p.band_lo) * rb;
                char *dst = f.lattice[k].base + (size_t)(d == 0 ? lo - rows : hi) * rb;
                a.job[a.njobs++] = PullJob{src, dst, bytes};
                total += bytes;
            }
        }
        if (with_state) {
            w.wait[w.nwait++] = FlagRef{p.flags + kFlagReadyState * kFlagStride, b->ticket};
            a.signal[a.nsignal++] = FlagRef{p.flags + (kFlagPulledState + mine_at_peer) * kFlagStride, b->ticket};
            const char *src = p.staging + (size_t)mine_at_peer * b->staging_side_bytes;   // the upper neighbour's BOTTOM rows, the lower one's TOP rows
            for (int k = 0; k < 3; k++) {
                const size_t rb = f.state[k].row_bytes, bytes = (size_t)SVGF_BAND_APRON * rb;
                char *dst = f.state[k].base + (size_t)(d == 0 ? lo - SVGF_BAND_APRON : hi) * rb;
                a.job[a.njobs++] = PullJob{src, dst, bytes};
                src += bytes;
                total += bytes;
            }
        }
    }
    BAND_TRY(launch_flags(b, w, stream));
    int grid = (int)(total / (256ull * 16 * 8)) + 1;
    if (grid > 64) grid = 64;
    band_pull_kernel<<<grid, 256, 0, stream>>>(a);
    b->extra_launches++;
    return band_cuda(b, cudaGetLastError());
}

}  // namespace

extern "C" {

// The a-trous levels 1..levels-1 of one band's frame as a list of steps (level 0 always covers the whole local image).
//  * a level below 3 recomputes its halo: it produces the band plus the rows the not-exchanging levels above it still
//    need (8 rows around the band for level 1 when level 2 follows), rounded out to whole row blocks of 12 * 2^level rows;
//  * a level >= 3 waits for its halo (2 * 2^level rows of the previous level's output from each neighbour) - before its
//    first launch, or, when it feeds no exchange itself, after the row blocks that do not read apron rows;
//  * a level whose successor is >= 3 runs the row blocks holding the rows its neighbours need FIRST, then the exchange
//    of those rows is posted, then the interior row blocks run.
int svgf_band_plan(int rank, int world, int band_lo, int band_hi, int local_rows, int levels, svgf_band_step *steps, int max_steps) {
    if (levels < 2 || levels > 5 || world < 1 || rank < 0 || rank >= world || band_lo < 0 || band_hi <= band_lo || band_hi > local_rows || !steps)
        return -1;
    const int first_exchanged = 3;
    int n = 0;
    auto push = [&](int kind, int level, int yb0, int nyb, int rows, int yb1 = 0, int nyb1 = 0) {
        if (n < max_steps) steps[n] = svgf_band_step{kind, level, yb0, nyb, rows, yb1, nyb1};
        n++;
    };
    for (int l = 1; l < levels; l++) {
        const int B = 12 << l;                     // rows per row block of this level's tile grid
        int reach = 0;
        for (int j = l + 1; j < levels && j < first_exchanged; j++) reach += 2 << j;
        const int r0 = band_lo - reach > 0 ? band_lo - reach : 0, r1 = band_hi + reach < local_rows ? band_hi + reach : local_rows;
        const int ybA = r0 / B, ybB = (r1 + B - 1) / B;
        const bool feeds_exchange = (l + 1 < levels) && (l + 1 >= first_exchanged) && world > 1;
        const bool waits = l >= first_exchanged && world > 1;
        if (waits && !feeds_exchange) {
            // a level that only CONSUMES a halo: the row blocks out of the halo's reach run first, under the exchange
            const int reach_in = 2 << l;
            int t1 = rank > 0 ? (band_lo + reach_in + B - 1) / B : ybA;       // [ybA, t1) read the top apron
            int b0 = rank + 1 < world ? (band_hi - reach_in) / B : ybB;       // [b0, ybB) read the bottom apron
            if (t1 > ybB) t1 = ybB;
            if (b0 < ybA) b0 = ybA;
            if (t1 < b0) {
                push(SVGF_BAND_STEP_LAUNCH, l, t1, b0 - t1, 0);
                push(SVGF_BAND_STEP_WAIT_HALO, l, 0, 0, reach_in);
                if (t1 > ybA && ybB > b0) push(SVGF_BAND_STEP_LAUNCH, l, ybA, t1 - ybA, 0, b0, ybB - b0);
                else if (t1 > ybA) push(SVGF_BAND_STEP_LAUNCH, l, ybA, t1 - ybA, 0);
                else if (ybB > b0) push(SVGF_BAND_STEP_LAUNCH, l, b0, ybB - b0, 0);
                continue;
            }
        }
        if (waits) push(SVGF_BAND_STEP_WAIT_HALO, l, 0, 0, 2 << l);
        if (!feeds_exchange) {
            push(SVGF_BAND_STEP_LAUNCH, l, ybA, ybB - ybA, 0);
            continue;
        }
        const int halo = 2 << (l + 1);
        int t1 = rank > 0 ? (band_lo + halo + B - 1) / B : ybA;              // top boundary blocks [ybA, t1)
        int b0 = rank + 1 < world ? (band_hi - halo) / B : ybB;              // bottom boundary blocks [b0, ybB)
        if (t1 > ybB) t1 = ybB;
        if (b0 < ybA) b0 = ybA;
        if (t1 >= b0) {                            // short band: the boundary blocks meet, nothing is left to overlap with
            push(SVGF_BAND_STEP_LAUNCH, l, ybA, ybB - ybA, 0);
            push(SVGF_BAND_STEP_EXCHANGE, l, 0, 0, halo);
        } else {
            // both boundary strips in ONE launch (two row-block ranges): a strip alone is one or two waves of tiles
            if (t1 > ybA && ybB > b0) push(SVGF_BAND_STEP_LAUNCH, l, ybA, t1 - ybA, 0, b0, ybB - b0);
            else if (t1 > ybA) push(SVGF_BAND_STEP_LAUNCH, l, ybA, t1 - ybA, 0);
            else if (ybB > b0) push(SVGF_BAND_STEP_LAUNCH, l, b0, ybB - b0, 0);
            push(SVGF_BAND_STEP_EXCHANGE, l, 0, 0, halo);
            push(SVGF_BAND_STEP_LAUNCH, l, t1, b0 - t1, 0);
        }
    }
    return n <= max_steps ? n : -1;
}

svgf_status svgf_band_unique_id(void *id_out) {
    if (!id_out) return SVGF_INVALID_ARG;
    if (!nccl().ok) return SVGF_UNSUPPORTED;
    NcclUniqueId id;
    if (nccl().GetUniqueId(&id) != 0) return SVGF_CUDA_ERROR;
    std::memcpy(id_out, &id, sizeof(id));
    return SVGF_OK;
}

enum BandTransport { kTransportNccl = 0, kTransportGroup = 1, kTransportIpc = 2 };

static svgf_status band_create(svgf_band **out, int device, int rank, int world, int width, int full_height, svgf_storage storage,
                               const void *unique_id, const int32_t *row_bounds, int transport) {
    const bool in_group = transport == kTransportGroup, use_nccl = transport == kTransportNccl;
    if (!out || world < 1 || rank < 0 || rank >= world || width <= 0 || full_height <= 0) return SVGF_INVALID_ARG;
    if (world > 1 && !unique_id && use_nccl) return SVGF_INVALID_ARG;
    *out = nullptr;
    int y0, y1;
    if (row_bounds) {
        if (row_bounds[0] != 0 || row_bounds[world] != full_height) return SVGF_INVALID_ARG;
        for (int g = 0; g < world; g++)
            if (row_bounds[g + 1] <= row_bounds[g]) return SVGF_INVALID_ARG;
        y0 = row_bounds[rank]; y1 = row_bounds[rank + 1];
        if (world > 1)
            for (int g = 0; g < world; g++)
                if (row_bounds[g + 1] - row_bounds[g] < SVGF_BAND_APRON) return SVGF_UNSUPPORTED;   // a band must fill its neighbours' aprons
    } else {
        const int base = full_height / world, rem = full_height % world;
        y0 = rank * base + (rank < rem ? rank : rem);
        y1 = y0 + base + (rank < rem ? 1 : 0);
        if (world > 1 && base < SVGF_BAND_APRON) return SVGF_UNSUPPORTED;
    }
    if (world > 1 && use_nccl && !nccl().ok) return SVGF_UNSUPPORTED;
    svgf_band *b = new (std::nothrow) svgf_band();
    if (!b) return SVGF_CUDA_ERROR;
    b->device = device; b->rank = rank; b->world = world; b->W = width; b->H = full_height;
    b->in_group = in_group;
    b->ipc = transport == kTransportIpc;
    b->y0 = y0; b->y1 = y1;
    b->ly0 = (world > 1 && y0 - SVGF_BAND_APRON > 0) ? y0 - SVGF_BAND_APRON : (world > 1 ? 0 : y0);
    b->ly1 = (world > 1 && y1 + SVGF_BAND_APRON < full_height) ? y1 + SVGF_BAND_APRON : (world > 1 ? full_height : y1);
    svgf_status st = svgf_create(&b->ctx, device, width, b->local_rows(), storage);
    if (st != SVGF_OK) { delete b; return st; }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) {
        // highest priority: the level kernels fill every SM (two CTAs take all of its shared memory), so the exchange's
        // CTAs only get in as tiles retire - and must then be picked ahead of the thousands of tiles still queued
        int lowest = 0, greatest = 0;
        e = cudaDeviceGetStreamPriorityRange(&lowest, &greatest);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&b->side, cudaStreamNonBlocking, greatest);
    }
    cudaEvent_t *evs[] = {&b->ev_l0, &b->ev_boundary[0], &b->ev_boundary[1], &b->ev_halo[0], &b->ev_halo[1], &b->ev_state, &b->ev_pulled};
    for (cudaEvent_t *ev : evs)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    int nr = 0;
    if (e == cudaSuccess && world > 1 && b->ipc) {
        // exported memory: plain cudaMalloc (cudaIpcGetMemHandle does not take pooled or virtual-memory allocations)
        const size_t ct = storage == SVGF_STORE_F32 ? 16 : 8, mt = storage == SVGF_STORE_F32 ? 8 : 4;
        b->staging_side_bytes = ((size_t)SVGF_BAND_APRON * width * (ct + mt + 1) + 255) & ~(size_t)255;
        e = cudaMalloc(&b->staging, 2 * b->staging_side_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&b->flags, (kFlagWords + 2) * kFlagStride * sizeof(unsigned));
        if (e == cudaSuccess) e = cudaMemset(b->flags, 0, (kFlagWords + 2) * kFlagStride * sizeof(unsigned));
        if (e == cudaSuccess) {
            b->pull_counter = b->flags + kFlagWords * kFlagStride;
            b->pull_counter_state = b->flags + (kFlagWords + 1) * kFlagStride;
            int lowest = 0, greatest = 0;
            e = cudaDeviceGetStreamPriorityRange(&lowest, &greatest);
            if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&b->side_state, cudaStreamNonBlocking, greatest);
        }
        if (e == cudaSuccess && svgf::lattice_prepare(b->ctx, nullptr) != SVGF_OK) e = cudaErrorMemoryAllocation;
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    if (e == cudaSuccess && world > 1 && use_nccl) {
        NcclUniqueId id;
        std::memcpy(&id, unique_id, sizeof(id));
        nr = nccl().CommInitRank(&b->comm, world, id, rank);
    }
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess || nr != 0) { svgf_band_destroy(b); return SVGF_CUDA_ERROR; }
    *out = b;
    return SVGF_OK;
}

svgf_status svgf_band_create(svgf_band **out, int device, int rank, int world, int width, int full_height, svgf_storage storage,
                             const void *unique_id, const int32_t *row_bounds) {
    return band_create(out, device, rank, world, width, full_height, storage, unique_id, row_bounds, kTransportNccl);
}

svgf_status svgf_band_create_ipc(svgf_band **out, int device, int rank, int world, int width, int full_height, svgf_storage storage,
                                 const int32_t *row_bounds) {
    return band_create(out, device, rank, world, width, full_height, storage, nullptr, row_bounds, kTransportIpc);
}

svgf_status svgf_band_ipc_export(svgf_band *b, void *blob_out) {
    if (!b || !blob_out || !b->ipc) return SVGF_INVALID_ARG;
    IpcBlob blob;
    std::memset(&blob, 0, sizeof(blob));
    if (b->world > 1) {
        const svgf_ctx *c = b->ctx;
        blob.magic = kIpcMagic; blob.width = b->W; blob.storage = (int)c->storage;
        blob.band_lo = b->band_lo(); blob.band_hi = b->band_hi(); blob.local_rows = b->local_rows(); blob.pitch_pairs = c->lat.pitch_pairs;
        DeviceScope dev;
        BAND_TRY(band_cuda(b, dev.enter(b->device)));
        void *planes[6] = {c->lat.sc[0].c0, c->lat.sc[0].c1, c->lat.sc[0].lz, c->lat.sc[1].c0, c->lat.sc[1].c1, c->lat.sc[1].lz};
        for (int k = 0; k < 6; k++) BAND_TRY(band_cuda(b, cudaIpcGetMemHandle(&blob.planes[k], planes[k])));
        BAND_TRY(band_cuda(b, cudaIpcGetMemHandle(&blob.staging, b->staging)));
        BAND_TRY(band_cuda(b, cudaIpcGetMemHandle(&blob.flags, b->flags)));
    }
    std::memset(blob_out, 0, SVGF_BAND_IPC_BYTES);
    std::memcpy(blob_out, &blob, sizeof(blob));
    return SVGF_OK;
}

svgf_status svgf_band_ipc_connect(svgf_band *b, const void *up_blob, const void *down_blob) {
    if (!b || !b->ipc || b->ipc_connected) return SVGF_INVALID_ARG;
    if ((b->rank > 0) != (up_blob != nullptr) || (b->rank + 1 < b->world) != (down_blob != nullptr)) return SVGF_INVALID_ARG;
    DeviceScope dev;
    BAND_TRY(band_cuda(b, dev.enter(b->device)));
    const void *blobs[2] = {up_blob, down_blob};
    for (int d = 0; d < 2; d++) {
        if (!blobs[d]) continue;
        IpcBlob blob;
        std::memcpy(&blob, blobs[d], sizeof(blob));
        if (blob.magic != kIpcMagic || blob.width != b->W || blob.storage != (int)b->ctx->storage || blob.pitch_pairs != b->ctx->lat.pitch_pairs)
            return SVGF_INVALID_ARG;
        IpcPeer &p = b->peer[d];
        p.band_lo = blob.band_lo; p.band_hi = blob.band_hi;
        for (int k = 0; k < 6; k++)
            BAND_TRY(band_cuda(b, cudaIpcOpenMemHandle((void **)&p.planes[k], blob.planes[k], cudaIpcMemLazyEnablePeerAccess)));
        BAND_TRY(band_cuda(b, cudaIpcOpenMemHandle((void **)&p.staging, blob.staging, cudaIpcMemLazyEnablePeerAccess)));
        BAND_TRY(band_cuda(b, cudaIpcOpenMemHandle((void **)&p.flags, blob.flags, cudaIpcMemLazyEnablePeerAccess)));
        p.open = true;
    }
    b->ipc_connected = true;
    return SVGF_OK;
}

void svgf_band_destroy(svgf_band *b) {
    if (!b) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(b->device);
    if (b->side) cudaStreamSynchronize(b->side);
    if (b->side_state) { cudaStreamSynchronize(b->side_state); cudaStreamDestroy(b->side_state); }
    if (b->up) { cudaStreamSynchronize(b->up->side); b->up->down = nullptr; }       // a neighbour's copies read this band's planes
    if (b->down) { cudaStreamSynchronize(b->down->side); b->down->up = nullptr; }
    if (b->comm) nccl().CommDestroy(b->comm);
    for (IpcPeer &p : b->peer) {
        for (char *q : p.planes)
            if (q) cudaIpcCloseMemHandle(q);
        if (p.staging) cudaIpcCloseMemHandle(p.staging);
        if (p.flags) cudaIpcCloseMemHandle(p.flags);
    }
    if (b->staging) cudaFree(b->staging);
    if (b->flags) cudaFree(b->flags);
    if (b->side) cudaStreamDestroy(b->side);
    cudaEvent_t evs[] = {b->ev_l0, b->ev_boundary[0], b->ev_boundary[1], b->ev_halo[0], b->ev_halo[1], b->ev_state, b->ev_pulled};
    for (cudaEvent_t ev : evs)
        if (ev) cudaEventDestroy(ev);
    svgf_destroy(b->ctx);
    if (prev >= 0 && prev != b->device) cudaSetDevice(prev);
    delete b;
}

void svgf_band_rows(const svgf_band *b, int32_t rows[4]) {
    if (!b || !rows) return;
    rows[0] = b->y0; rows[1] = b->y1; rows[2] = b->ly0; rows[3] = b->ly1;
}

uint64_t svgf_band_launch_count(const svgf_band *b) { return b ? svgf_launch_count(b->ctx) + b->extra_launches : 0; }
int svgf_band_last_error(const svgf_band *b) { return !b ? 0 : (b->last_err ? b->last_err : svgf_last_cuda_error(b->ctx)); }

svgf_status svgf_band_sync(svgf_band *b, void *stream) {
    if (!b) return SVGF_INVALID_ARG;
    if (b->state_pending) {
        BAND_TRY(band_cuda(b, cudaStreamWaitEvent((cudaStream_t)stream, b->ev_state, 0)));
        b->state_pending = false;
    }
    return SVGF_OK;
}

svgf_status svgf_band_reset(svgf_band *b, const svgf_frame_buffers *bufs, void *stream) {
    if (!b) return SVGF_INVALID_ARG;
    BAND_TRY(svgf_band_sync(b, stream));
    return svgf_reset(b->ctx, bufs, stream);
}

namespace {

// Everything of the frame up to and including level 0, and the plan of the remaining levels.
svgf_status frame_begin(svgf_band *b, const svgf_params *params, const svgf_gbuffer gbuf[2], const svgf_frame_buffers *bufs, void *stream) {
    svgf_ctx *c = b->ctx;
    svgf_band::Frame &f = b->f;
    const int N = params->atrous_iterations;
    if (N < 2 || N > 5) return SVGF_UNSUPPORTED;
    if (bufs->ping_pong != 0 && bufs->ping_pong != 1) return SVGF_INVALID_ARG;
    const int P = bufs->ping_pong;
    f.params = *params; f.bufs = *bufs; f.s = (cudaStream_t)stream; f.N = N;
    cudaStream_t s = f.s;
    b->dry_run = (params->flags & SVGF_FLAG_BAND_NO_EXCHANGE) != 0;

    // previous-frame state in the aprons: posted under the previous frame's levels
    BAND_TRY(svgf_band_sync(b, stream));
    if (b->ipc) {
        if (!b->ipc_connected) return SVGF_INVALID_ARG;
        b->ticket++;
        if (!b->dry_run) {
            // off the main stream: the check runs beside the temporal pass; level 0 (the first kernel that overwrites rows a
            // neighbour reads) and the staging copies (same side stream) are ordered behind it
            BAND_TRY(ipc_wait_pulled(b, b->side));
            BAND_TRY(band_cuda(b, cudaEventRecord(b->ev_pulled, b->side)));
        }
    }
    if (b->in_group) {
        // a neighbour copies rows out of THIS band's planes on its own side stream: this frame must not overwrite them
        // before those copies (all posted during the previous group frame) have run
        for (svgf_band *nb : {b->up, b->down}) {
            if (!nb) continue;
            for (cudaEvent_t ev : {nb->ev_halo[0], nb->ev_halo[1], nb->ev_state}) BAND_TRY(band_cuda(b, cudaStreamWaitEvent(s, ev, 0)));
        }
    }

    // temporal + variance over the whole local image = svgf_frame with no a-trous level (variance output in filter[0])
    svgf_params p0 = *params;
    p0.atrous_iterations = 0;
    BAND_TRY(svgf_frame(c, &p0, gbuf, bufs, stream));
    if (!svgf::staged_run_possible(c, params, bufs->filter[0], bufs->filter[1], bufs->render[P], 0, N)) return SVGF_UNSUPPORTED;
    BAND_TRY(svgf::lattice_prepare(c, s));
    f.slot = c->guide_cur;
    c->dispatch_n = 0;

    // level 0 over the whole local image: it also writes the normal planes every later level reads in the apron, and the
    // colour history of the apron rows is replaced by the neighbours' below
    if (b->ipc && !b->dry_run) BAND_TRY(band_cuda(b, cudaStreamWaitEvent(s, b->ev_pulled, 0)));
    BAND_TRY(svgf::staged_level(c, params, f.slot, 0, 0, bufs->filter[0], 0, nullptr, bufs->render[P], 0, 0, 0, 0, false, s));
    // next frame's previous-frame state - colour history (level 0's output), moments and history lengths - is final from here
    // on.  It travels with the LAST halo exchange of the frame (one NCCL launch fewer; it is not needed before the next
    // frame's temporal pass), or on its own right away when no level exchanges a halo.
    BAND_TRY(band_cuda(b, cudaEventRecord(b->ev_l0, s)));
    const size_t ct = c->storage == SVGF_STORE_F32 ? 16 : 8, mt = c->storage == SVGF_STORE_F32 ? 8 : 4;
    f.state[0] = Plane{(char *)bufs->render[P], (size_t)b->W * ct};
    f.state[1] = Plane{(char *)bufs->moments[P], (size_t)b->W * mt};
    f.state[2] = Plane{(char *)bufs->history, (size_t)b->W};
    if (b->ipc) {
        // peer-memory transport: the state rows are published and the neighbours' pulled right away, by a kernel of their own
        // under levels 1-2 - a separate launch costs nothing here, and the last halo exchange (the exposed one) stays small
        // on a stream of their own: a neighbour whose level 0 ends later must not hold up this band's halo exchanges
        BAND_TRY(band_cuda(b, cudaStreamWaitEvent(b->side_state, b->ev_l0, 0)));
        if (!b->dry_run) {
            BAND_TRY(band_cuda(b, cudaStreamWaitEvent(b->side_state, b->ev_pulled, 0)));   // the staging block has been fetched
            BAND_TRY(ipc_publish_state(b));
            b->state_ticket = b->ticket;
        }
        BAND_TRY(exchange_ipc(b, false, true));
        BAND_TRY(band_cuda(b, cudaEventRecord(b->ev_state, b->side_state)));
        b->state_pending = true;
    }

    // levels 1..N-1 as planned by svgf_band_plan (the same function the CPU tests check)
    f.n_steps = svgf_band_plan(b->rank, b->world, b->band_lo(), b->band_hi(), b->local_rows(), N, f.steps, 32);
    if (f.n_steps < 0) return SVGF_UNSUPPORTED;
    f.last_exchange = -1;
    for (int i = 0; i < f.n_steps; i++)
        if (f.steps[i].kind == SVGF_BAND_STEP_EXCHANGE) f.last_exchange = i;
    f.src = 0;                                     // lattice colour set holding the input of the current level
    f.cur_level = 1; f.n_halo = 0;
    f.pc = (f.last_exchange < 0 && !b->ipc) ? -1 : 0;   // -1: the state-only exchange comes first (NCCL / group transports)
    return SVGF_OK;
}

// Runs the frame's steps up to the next exchange.  *stopped = true: f.sets / f.ready / f.done describe the exchange to
// perform (frame_exchanged() then continues); false: the frame has been issued completely.
svgf_status frame_run(svgf_band *b, bool *stopped) {
    svgf_ctx *c = b->ctx;
    svgf_band::Frame &f = b->f;
    cudaStream_t s = f.s;
    *stopped = false;
    if (f.pc < 0) {                                // no level exchanges a halo: the state travels on its own
        f.sets[0] = PlaneSet{f.state, 3, SVGF_BAND_APRON};
        f.n_sets = 1; f.ready = b->ev_l0; f.done = nullptr; f.carries_state = true;
        *stopped = true;
        return SVGF_OK;
    }
    for (; f.pc < f.n_steps; f.pc++) {
        const svgf_band_step &st = f.steps[f.pc];
        if (st.level != f.cur_level) { f.src = 1 - f.src; f.cur_level = st.level; }
        const bool last = (st.level == f.N - 1);
        if (st.kind == SVGF_BAND_STEP_WAIT_HALO) {
            BAND_TRY(band_cuda(b, cudaStreamWaitEvent(s, b->ev_halo[(st.level - 3) & 1], 0)));
        } else if (st.kind == SVGF_BAND_STEP_LAUNCH) {
            BAND_TRY(svgf::staged_level(c, &f.params, f.slot, st.level, last ? 2 : 1, nullptr, f.src, last ? f.bufs.filter[0] : nullptr, nullptr,
                                        st.yblock0, st.nyblocks, st.yblock1, st.nyblocks1, false, s));
        } else {   // SVGF_BAND_STEP_EXCHANGE: rows of THIS level's output, for level + 1
            cudaEvent_t evb = b->ev_boundary[f.n_halo & 1];
            BAND_TRY(band_cuda(b, cudaEventRecord(evb, s)));                   // (implies level 0: same stream, earlier)
            const svgf::LatticeColour &dst = c->lat.sc[1 - f.src];
            const size_t row_bytes = (size_t)c->lat.pitch_pairs * 16, pad = (size_t)svgf::kLatPadY * row_bytes;
            f.lattice[0] = Plane{(char *)dst.c0 + pad, row_bytes};
            f.lattice[1] = Plane{(char *)dst.c1 + pad, row_bytes};
            f.lattice[2] = Plane{(char *)dst.lz + pad, row_bytes};
            f.sets[0] = PlaneSet{f.lattice, 3, st.rows};
            f.sets[1] = PlaneSet{f.state, 3, SVGF_BAND_APRON};
            f.carries_state = (f.pc == f.last_exchange) && !b->ipc;
            f.n_sets = f.carries_state ? 2 : 1;
            f.ready = evb; f.done = b->ev_halo[(st.level + 1 - 3) & 1];
            *stopped = true;
            return SVGF_OK;
        }
    }
    return SVGF_OK;
}

// The exchange frame_run() stopped at, on the side stream; then the bookkeeping that lets the frame continue.
svgf_status frame_exchange(svgf_band *b) {
    svgf_band::Frame &f = b->f;
    BAND_TRY(band_cuda(b, cudaStreamWaitEvent(b->side, f.ready, 0)));
    if (b->in_group) {
        for (svgf_band *nb : {b->up, b->down})
            if (nb) BAND_TRY(band_cuda(b, cudaStreamWaitEvent(b->side, nb->f.ready, 0)));
        BAND_TRY(exchange_copy(b));
    } else if (b->ipc) {
        BAND_TRY(exchange_ipc(b, true, false));
    } else {
        BAND_TRY(exchange(b, f.sets, f.n_sets));
    }
    if (f.done) BAND_TRY(band_cuda(b, cudaEventRecord(f.done, b->side)));
    if (f.carries_state) {
        BAND_TRY(band_cuda(b, cudaEventRecord(b->ev_state, b->side)));
        b->state_pending = true;
    }
    if (f.pc >= 0) f.n_halo++;
    f.pc++;                                        // -1 -> 0: the steps follow the state-only exchange
    return SVGF_OK;
}

}  // namespace

svgf_status svgf_band_frame(svgf_band *b, const svgf_params *params, const svgf_gbuffer gbuf[2], const svgf_frame_buffers *bufs,
                            void *stream) {
    if (!b || !params || !gbuf || !bufs) return SVGF_INVALID_ARG;
    if (b->in_group) return SVGF_INVALID_ARG;                                  // a group's bands advance together: svgf_band_group_frame
    if (b->world == 1) return svgf_frame(b->ctx, params, gbuf, bufs, stream);  // one band = the whole frame
    DeviceScope dev;
    BAND_TRY(band_cuda(b, dev.enter(b->device)));
    BAND_TRY(frame_begin(b, params, gbuf, bufs, stream));
    for (;;) {
        bool stopped = false;
        BAND_TRY(frame_run(b, &stopped));
        if (!stopped) return SVGF_OK;
        BAND_TRY(frame_exchange(b));
    }
}

svgf_status svgf_band_create_group(svgf_band **out, const int32_t *devices, int world, int width, int full_height, svgf_storage storage,
                                   const int32_t *row_bounds) {
    if (!out || !devices || world < 1) return SVGF_INVALID_ARG;
    for (int g = 0; g < world; g++) out[g] = nullptr;
    for (int g = 0; g < world; g++) {
        const svgf_status st = band_create(&out[g], devices[g], g, world, width, full_height, storage, nullptr, row_bounds, kTransportGroup);
        if (st != SVGF_OK) {
            for (int k = 0; k < g; k++) { svgf_band_destroy(out[k]); out[k] = nullptr; }
            return st;
        }
    }
    for (int g = 0; g < world; g++) {
        out[g]->up = g > 0 ? out[g - 1] : nullptr;
        out[g]->down = g + 1 < world ? out[g + 1] : nullptr;
    }
    // neighbours on different devices: direct peer copies where the hardware allows (otherwise the runtime stages them)
    int prev = -1;
    cudaGetDevice(&prev);
    for (int g = 0; g + 1 < world; g++) {
        const int a = devices[g], b = devices[g + 1];
        if (a == b) continue;
        int ok = 0;
        if (cudaDeviceCanAccessPeer(&ok, a, b) == cudaSuccess && ok && cudaSetDevice(a) == cudaSuccess) cudaDeviceEnablePeerAccess(b, 0);
        if (cudaDeviceCanAccessPeer(&ok, b, a) == cudaSuccess && ok && cudaSetDevice(b) == cudaSuccess) cudaDeviceEnablePeerAccess(a, 0);
        cudaGetLastError();                    // already enabled is fine
    }
    if (prev >= 0) cudaSetDevice(prev);
    return SVGF_OK;
}

svgf_status svgf_band_group_frame(svgf_band *const *bands, int world, const svgf_params *params, const svgf_gbuffer *gbufs,
                                  const svgf_frame_buffers *bufs, void *const *streams) {
    if (!bands || !params || !gbufs || !bufs || !streams || world < 1) return SVGF_INVALID_ARG;
    for (int g = 0; g < world; g++)
        if (!bands[g] || !bands[g]->in_group || bands[g]->world != world || bands[g]->rank != g) return SVGF_INVALID_ARG;
    if (world == 1) return svgf_frame(bands[0]->ctx, params, gbufs, bufs, streams[0]);
    // every band up to its next exchange, then every band's exchange (each waits for its neighbours' `ready` events, which
    // by then have all been recorded), and so on: all bands stop at the same exchanges in the same order
    for (int g = 0; g < world; g++) {
        DeviceScope dev;
        BAND_TRY(band_cuda(bands[g], dev.enter(bands[g]->device)));
        BAND_TRY(frame_begin(bands[g], params, gbufs + 2 * g, bufs + g, streams[g]));
    }
    for (;;) {
        int n_stopped = 0;
        for (int g = 0; g < world; g++) {
            DeviceScope dev;
            BAND_TRY(band_cuda(bands[g], dev.enter(bands[g]->device)));
            bool stopped = false;
            BAND_TRY(frame_run(bands[g], &stopped));
            n_stopped += stopped ? 1 : 0;
        }
        if (n_stopped == 0) return SVGF_OK;
        if (n_stopped != world) return SVGF_UNSUPPORTED;                       // cannot happen: the plans exchange alike
        for (int g = 0; g < world; g++) {
            DeviceScope dev;
            BAND_TRY(band_cuda(bands[g], dev.enter(bands[g]->device)));
            BAND_TRY(frame_exchange(bands[g]));
        }
    }
}

}  // extern "C"
