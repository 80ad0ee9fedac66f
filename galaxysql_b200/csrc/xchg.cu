// xchg.cu — hash-partition exchange behind gsql_xchg_* and gsql_comm_*.
//
// Reference path replaced (EX/ = polardbx-executor/src/main/java/com/alibaba/polardbx/executor/):
//   local : EX/mpp/operator/PartitioningExchanger.java:71-135 (consumeChunk: per-row partition id, per-destination
//           chunk build), HashBucketFunction (PartitionedOutputCollector.java:282-302)
//   remote: EX/mpp/operator/PartitionedOutputCollector.java:170-196 (partitionPage) + HashPartitionFunction:253-279,
//           PagesSerde / PartitionedOutputBuffer / ExchangeClient HTTP pull (ExchangeClient.java:62-548)
// Destination ids are bit-exact with ExecUtils.partition(Chunk.hashCode) so GPU and stock Java tasks route the
// same row to the same consumer.
//
// B200 shape: histogram (keys only) -> exclusive scan over [partition][block] -> scatter of all columns into one
// buffer whose partitions are contiguous; across GPUs the contiguous segments are exchanged with one grouped
// ncclSend/ncclRecv AllToAllv per column over NVLink 5 / NVSwitch — no serialisation, no compression.
// NCCL is bound at run time (dlopen) so that the library shares the process's NCCL with torch.distributed.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace {

constexpr int XBLOCK = 256;

struct XParams {
    DColSet in;
    KeySet keys;
    int32_t nparts, pow2;
    int64_t rows, chunk;  // rows per block
    int32_t nblocks, pad;
};

struct XOut {
    void *data[GSQL_MAX_COLS];
    uint8_t *nulls[GSQL_MAX_COLS];
};

__device__ __forceinline__ int row_part(const XParams &P, int64_t r) {
    return gsql_partition_of(gsql_row_hash(P.keys, r), P.nparts, P.pow2 != 0);
}

// hist[p * nblocks + b] = rows of block b's chunk routed to p
__global__ void __launch_bounds__(XBLOCK) k_xchg_hist(const __grid_constant__ XParams P, int64_t *__restrict__ hist) {
    extern __shared__ unsigned int sh[];
    for (int i = threadIdx.x; i < P.nparts; i += XBLOCK) sh[i] = 0;
    __syncthreads();
    int64_t r0 = (int64_t)blockIdx.x * P.chunk;
    int64_t r1 = r0 + P.chunk < P.rows ? r0 + P.chunk : P.rows;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += XBLOCK) atomicAdd(&sh[row_part(P, r)], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < P.nparts; i += XBLOCK) hist[(int64_t)i * P.nblocks + blockIdx.x] = sh[i];
}

__global__ void __launch_bounds__(XBLOCK)
    k_xchg_scatter(const __grid_constant__ XParams P, const int64_t *__restrict__ offs, const __grid_constant__ XOut O) {
    extern __shared__ unsigned long long cur[];
    for (int i = threadIdx.x; i < P.nparts; i += XBLOCK) cur[i] = (unsigned long long)offs[(int64_t)i * P.nblocks + blockIdx.x];
    __syncthreads();
    int64_t r0 = (int64_t)blockIdx.x * P.chunk;
    int64_t r1 = r0 + P.chunk < P.rows ? r0 + P.chunk : P.rows;
    for (int64_t base = r0; base < r1; base += XBLOCK) {
        int64_t r = base + threadIdx.x;
        bool live = r < r1;
        int p = live ? row_part(P, r) : -1;
        // warp-aggregated cursor bump: one shared-memory atomic per distinct destination per warp
        unsigned peers = __match_any_sync(0xffffffffu, p);
        int lane = threadIdx.x & 31;
        int leader = __ffs(peers) - 1;
        unsigned long long basepos = 0;
        if (live && lane == leader) basepos = atomicAdd(&cur[p], (unsigned long long)__popc(peers));
        basepos = __shfl_sync(0xffffffffu, basepos, leader);
        if (!live) continue;
        int64_t pos = (int64_t)basepos + __popc(peers & ((1u << lane) - 1));
        for (int c = 0; c < P.in.n; c++) {
            const DCol &col = P.in.c[c];
            if (col.type == GSQL_T_INT32) reinterpret_cast<int32_t *>(O.data[c])[pos] = reinterpret_cast<const int32_t *>(col.data)[r];
            else reinterpret_cast<int64_t *>(O.data[c])[pos] = reinterpret_cast<const int64_t *>(col.data)[r];
            if (O.nulls[c]) O.nulls[c][pos] = col.nulls ? col.nulls[r] : 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------ NCCL binding
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    // Prefer the NCCL already in the process (torch's bundled one), else the system library.
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return &api;
    api.handle = h;
#define BIND(field, sym) *(void **)(&api.field) = dlsym(h, sym)
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(CommSplit, "ncclCommSplit");
    BIND(AllGather, "ncclAllGather");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Send && api.Recv && api.GroupStart &&
             api.GroupEnd && api.GetErrorString;
    return &api;
}

#define GSQL_NCCL(ctx, call)                                                                                            \
    do {                                                                                                                \
        ncclResult_t _r = (call);                                                                                       \
        if (_r != ncclSuccess) {                                                                                        \
            (ctx)->sticky = true;                                                                                       \
            return gsql_set_error((ctx), GSQL_E_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call, nccl_api()->GetErrorString(_r)); \
        }                                                                                                               \
    } while (0)

int grid_rows(gsql_ctx *ctx, int64_t rows, int block, int per_sm) {
    int64_t g = div_up(rows, block);
    int64_t cap = (int64_t)ctx->sm_count * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

struct gsql_xchg {
    gsql_ctx *ctx;
    gsql_xchg_spec spec;
    // send staging of the shuffle (partition-ordered columns), kept across calls: a shuffle of N rows stages N rows,
    // and re-allocating those gigabytes on every call raced the stream-ordered frees of the previous one
    DevBuf sdata[GSQL_MAX_COLS], snull[GSQL_MAX_COLS];
};

extern "C" gsql_status gsql_xchg_create(gsql_ctx *ctx, const gsql_xchg_spec *spec, gsql_xchg **out) {
    if (!ctx || !spec || !out) return GSQL_E_INVALID;
    *out = nullptr;
    const gsql_xchg_spec &s = *spec;
    if (s.n_cols < 1 || s.n_cols > GSQL_MAX_COLS || s.n_channels < 0 || s.n_channels > GSQL_MAX_KEYS || s.nparts < 1 || s.nparts > GSQL_MAX_PARTS)
        return gsql_set_error(ctx, GSQL_E_INVALID, "bad exchange spec");
    for (int i = 0; i < s.n_cols; i++)
        if (s.types[i] < GSQL_T_INT32 || s.types[i] > GSQL_T_FP64) return gsql_set_error(ctx, GSQL_E_UNSUPPORTED, "column %d type", i);
    for (int i = 0; i < s.n_channels; i++)
        if (s.channels[i] < 0 || s.channels[i] >= s.n_cols || s.key_types[i] < GSQL_T_INT32 || s.key_types[i] > GSQL_T_FP64)
            return gsql_set_error(ctx, GSQL_E_INVALID, "channel %d", i);
    gsql_xchg *x = new gsql_xchg();
    x->ctx = ctx;
    gsql_ctx_retain(ctx);
    x->spec = s;
    *out = x;
    return GSQL_OK;
}

extern "C" void gsql_xchg_destroy(gsql_xchg *x) {
    if (!x) return;
    gsql_ctx *ctx = x->ctx;
    cudaSetDevice(ctx->device);
    delete x;
    if (!ctx->sticky) cudaStreamSynchronize(ctx->stream);  // stream-ordered frees have really happened (see gsql_join_destroy)
    gsql_ctx_release(ctx);
}

// Partitions staged device columns into `O` (device).  d_part_offsets receives nparts+1 int64 offsets (device).
static gsql_status partition_device(gsql_xchg *x, const StagedBatch &sb, const XOut &O, DevBuf *offs_out, int64_t *host_counts) {
    gsql_ctx *ctx = x->ctx;
    const gsql_xchg_spec &s = x->spec;
    XParams P;
    memset(&P, 0, sizeof(P));
    P.in.n = sb.ncols;
    for (int i = 0; i < sb.ncols; i++) P.in.c[i] = sb.cols[i];
    P.keys.n = s.n_channels;
    for (int i = 0; i < s.n_channels; i++) {
        P.keys.c[i] = sb.cols[s.channels[i]];
        P.keys.utype[i] = s.key_types[i];
    }
    P.nparts = s.nparts;
    P.pow2 = (s.nparts & -s.nparts) == s.nparts;
    P.rows = sb.rows;
    int nblocks = grid_rows(ctx, sb.rows, 4096, 8);
    P.chunk = div_up(sb.rows, nblocks);
    P.chunk = div_up(P.chunk, XBLOCK) * XBLOCK;
    nblocks = (int)div_up(sb.rows, P.chunk);
    if (nblocks < 1) nblocks = 1;
    P.nblocks = nblocks;
    int64_t nh = (int64_t)s.nparts * nblocks;
    DevBuf hist, tmp;
    GSQL_TRY(hist.alloc(ctx, (size_t)(nh + 1) * 8));
    GSQL_TRY(offs_out->alloc(ctx, (size_t)(nh + 1) * 8));
    GSQL_CUDA(ctx, cudaMemsetAsync((char *)hist.p + nh * 8, 0, 8, ctx->stream));
    {
        KernelScope ks(ctx, "xchg_hist");
        k_xchg_hist<<<nblocks, XBLOCK, s.nparts * sizeof(unsigned int), ctx->stream>>>(P, hist.as<int64_t>());
    }
    GSQL_CUDA(ctx, cudaGetLastError());
    size_t tb = 0;
    GSQL_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, hist.as<int64_t>(), offs_out->as<int64_t>(), nh + 1, ctx->stream));
    GSQL_TRY(tmp.alloc(ctx, tb));
    {
        KernelScope ks(ctx, "xchg_scan");
        GSQL_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, tb, hist.as<int64_t>(), offs_out->as<int64_t>(), nh + 1, ctx->stream));
    }
    {
        KernelScope ks(ctx, "xchg_scatter");
        k_xchg_scatter<<<nblocks, XBLOCK, s.nparts * sizeof(unsigned long long), ctx->stream>>>(P, offs_out->as<int64_t>(), O);
    }
    GSQL_CUDA(ctx, cudaGetLastError());
    if (host_counts) {  // partition p starts at offs[p * nblocks]
        std::vector<int64_t> starts((size_t)s.nparts + 1);
        GSQL_CUDA(ctx, cudaMemcpy2DAsync(starts.data(), 8, offs_out->p, (size_t)nblocks * 8, 8, (size_t)s.nparts, cudaMemcpyDeviceToHost, ctx->stream));
        GSQL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        starts[(size_t)s.nparts] = sb.rows;
        for (int p = 0; p < s.nparts; p++) host_counts[p] = starts[(size_t)p + 1] - starts[(size_t)p];
    }
    return GSQL_OK;
}

extern "C" gsql_status gsql_xchg_partition(gsql_xchg *x, const gsql_batch *in, gsql_batch *out, int64_t *part_counts) {
    if (!x || !in || !out || !part_counts) return GSQL_E_INVALID;
    gsql_ctx *ctx = x->ctx;
    if (ctx->sticky) return GSQL_E_CUDA;
    const gsql_xchg_spec &s = x->spec;
    GSQL_TRY(validate_batch(ctx, in, s.n_cols, s.types));
    GSQL_TRY(validate_batch(ctx, out, s.n_cols, s.types));
    if (in->mem != out->mem) return gsql_set_error(ctx, GSQL_E_INVALID, "in and out must live in the same memory space");
    for (int p = 0; p < s.nparts; p++) part_counts[p] = 0;
    out->rows = in->rows;
    if (in->rows == 0) return GSQL_OK;
    GSQL_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int c = 0; c < s.n_cols; c++)
        if (in->cols[c].nulls && !out->cols[c].nulls) return gsql_set_error(ctx, GSQL_E_INVALID, "output column %d needs a nulls buffer", c);
    StagedBatch sb;
    GSQL_TRY(stage_batch(ctx, in, &sb));
    XOut O;
    memset(&O, 0, sizeof(O));
    DevBuf odata[GSQL_MAX_COLS], onull[GSQL_MAX_COLS];
    for (int c = 0; c < s.n_cols; c++) {
        if (in->mem == GSQL_MEM_DEVICE) {
            O.data[c] = out->cols[c].data;
            O.nulls[c] = out->cols[c].nulls;
        } else {
            GSQL_TRY(odata[c].alloc(ctx, (size_t)in->rows * gsql_type_width(s.types[c])));
            O.data[c] = odata[c].p;
            if (out->cols[c].nulls) {
                GSQL_TRY(onull[c].alloc(ctx, (size_t)in->rows));
                O.nulls[c] = onull[c].as<uint8_t>();
            }
        }
    }
    DevBuf offs;
    GSQL_TRY(partition_device(x, sb, O, &offs, part_counts));
    if (in->mem == GSQL_MEM_HOST) {
        for (int c = 0; c < s.n_cols; c++) {
            GSQL_CUDA(ctx, cudaMemcpyAsync(out->cols[c].data, O.data[c], (size_t)in->rows * gsql_type_width(s.types[c]), cudaMemcpyDeviceToHost, ctx->stream));
            if (out->cols[c].nulls) GSQL_CUDA(ctx, cudaMemcpyAsync(out->cols[c].nulls, O.nulls[c], (size_t)in->rows, cudaMemcpyDeviceToHost, ctx->stream));
        }
        GSQL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return GSQL_OK;
}

// ------------------------------------------------------------------------------------------------ communicator
extern "C" gsql_status gsql_comm_unique_id(uint8_t id[128]) {
    NcclApi *api = nccl_api();
    if (!api->ok || !id) return GSQL_E_NCCL;
    ncclUniqueId uid;
    if (api->GetUniqueId(&uid) != ncclSuccess) return GSQL_E_NCCL;
    memcpy(id, uid.internal, 128);
    return GSQL_OK;
}

extern "C" gsql_status gsql_comm_init(gsql_ctx *ctx, int32_t nranks, int32_t rank, const uint8_t id[128]) {
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return GSQL_E_INVALID;
    NcclApi *api = nccl_api();
    if (!api->ok) return gsql_set_error(ctx, GSQL_E_NCCL, "libnccl.so.2 could not be loaded");
    GSQL_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    ncclComm_t comm;
    GSQL_NCCL(ctx, api->CommInitRank(&comm, nranks, uid, rank));
    ctx->nccl_comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    // stripe communicators
    int want = 3;
    if (const char *e = getenv("GSQL_XCHG_STRIPES")) want = atoi(e) - 1;
    if (want > 7) want = 7;
    ctx->n_extra = 0;
    if (nranks > 1 && api->CommSplit) {
        for (int i = 0; i < want; i++) {
            ncclComm_t extra = nullptr;
            if (api->CommSplit(comm, 0, rank, &extra, nullptr) != ncclSuccess || !extra) break;
            ctx->nccl_extra[i] = extra;
            cudaStreamCreateWithFlags(&ctx->xstreams[i], cudaStreamNonBlocking);
            ctx->n_extra = i + 1;
        }
    }
    return GSQL_OK;
}

extern "C" gsql_status gsql_comm_destroy(gsql_ctx *ctx) {
    if (!ctx) return GSQL_E_INVALID;
    if (ctx->nccl_comm) {
        cudaStreamSynchronize(ctx->stream);
        for (int i = 0; i < ctx->n_extra; i++) {
            cudaStreamSynchronize(ctx->xstreams[i]);
            nccl_api()->CommDestroy((ncclComm_t)ctx->nccl_extra[i]);
            cudaStreamDestroy(ctx->xstreams[i]);
            ctx->nccl_extra[i] = nullptr;
        }
        ctx->n_extra = 0;
        nccl_api()->CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    return GSQL_OK;
}

// ---- opt-in slabbed shuffle (GSQL_XCHG_SLABS=S > 1; parity-checked on one B200 rank at the end of r01, performance not measured yet) -----------
// The sequential path partitions the whole batch, then sends it: the scatter pass (HBM-bound) and the AllToAllv
// (NVLink-bound) never overlap.  Here the batch is cut into S row slabs.  All S histograms run first (keys only), one
// count exchange covers every (slab, destination) pair, and then slab i+1 is scattered on the context stream while the
// segments of slab i travel on the stripe streams.  Receive layout is the sequential path's: per source rank contiguous
// (its slabs one after the other).
static void fill_xparams(gsql_xchg *x, const StagedBatch &sb, int64_t row0, int64_t rows, XParams *Pp) {
    const gsql_xchg_spec &s = x->spec;
    XParams &P = *Pp;
    memset(&P, 0, sizeof(P));
    P.in.n = sb.ncols;
    for (int i = 0; i < sb.ncols; i++) {
        P.in.c[i] = sb.cols[i];
        P.in.c[i].data = (const char *)sb.cols[i].data + (size_t)row0 * gsql_type_width(sb.cols[i].type);
        if (sb.cols[i].nulls) P.in.c[i].nulls = sb.cols[i].nulls + row0;
    }
    P.keys.n = s.n_channels;
    for (int i = 0; i < s.n_channels; i++) {
        P.keys.c[i] = P.in.c[s.channels[i]];
        P.keys.utype[i] = s.key_types[i];
    }
    P.nparts = s.nparts;
    P.pow2 = (s.nparts & -s.nparts) == s.nparts;
    P.rows = rows;
    int nblocks = grid_rows(x->ctx, rows, 4096, 8);
    P.chunk = div_up(rows > 0 ? rows : 1, nblocks);
    P.chunk = div_up(P.chunk, XBLOCK) * XBLOCK;
    nblocks = (int)div_up(rows, P.chunk);
    if (nblocks < 1) nblocks = 1;
    P.nblocks = nblocks;
}

static gsql_status all_to_all_slabbed(gsql_xchg *x, const StagedBatch &sb, gsql_batch *out, int64_t out_capacity, int64_t *out_rows,
                                      int64_t *recv_counts, int S) {
    gsql_ctx *ctx = x->ctx;
    const gsql_xchg_spec &s = x->spec;
    NcclApi *api = nccl_api();
    const int R = ctx->nranks, K = ctx->n_extra;  // all NCCL traffic on the stripe streams: the context stream keeps scattering
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const int64_t n = sb.rows;
    const int64_t SR = div_up(div_up(n > 0 ? n : 1, S), XBLOCK) * XBLOCK;  // rows per slab (the last ones may be short or empty)
    DevBuf *sdata = x->sdata, *snull = x->snull;
    for (int c = 0; c < s.n_cols; c++) {
        GSQL_TRY(sdata[c].grow(ctx, (size_t)(n > 0 ? n : 1) * gsql_type_width(s.types[c]), 0));
        if (out->cols[c].nulls) GSQL_TRY(snull[c].grow(ctx, (size_t)(n > 0 ? n : 1), 0));
    }
    // ---- 1. histograms + scans of every slab
    std::vector<XParams> P((size_t)S);
    std::vector<DevBuf> hist((size_t)S), offs((size_t)S);
    std::vector<int64_t> rows_of((size_t)S, 0);
    DevBuf tmp;
    size_t tmp_bytes = 0;
    for (int i = 0; i < S; i++) {
        const int64_t lo = (int64_t)i * SR;
        rows_of[(size_t)i] = lo < n ? (n - lo < SR ? n - lo : SR) : 0;
        fill_xparams(x, sb, lo < n ? lo : 0, rows_of[(size_t)i], &P[(size_t)i]);
        const int64_t nh = (int64_t)s.nparts * P[(size_t)i].nblocks;
        GSQL_TRY(hist[(size_t)i].alloc(ctx, (size_t)(nh + 1) * 8));
        GSQL_TRY(offs[(size_t)i].alloc(ctx, (size_t)(nh + 1) * 8));
        GSQL_CUDA(ctx, cudaMemsetAsync(hist[(size_t)i].p, 0, (size_t)(nh + 1) * 8, ctx->stream));
        size_t tb = 0;
        GSQL_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, hist[(size_t)i].as<int64_t>(), offs[(size_t)i].as<int64_t>(), nh + 1, ctx->stream));
        if (tb > tmp_bytes) tmp_bytes = tb;
    }
    GSQL_TRY(tmp.alloc(ctx, tmp_bytes));
    for (int i = 0; i < S; i++) {
        const int64_t nh = (int64_t)s.nparts * P[(size_t)i].nblocks;
        if (rows_of[(size_t)i] > 0) {
            KernelScope ks(ctx, "xchg_hist");
            k_xchg_hist<<<P[(size_t)i].nblocks, XBLOCK, s.nparts * sizeof(unsigned int), ctx->stream>>>(P[(size_t)i], hist[(size_t)i].as<int64_t>());
        }
        size_t tb = tmp_bytes;
        GSQL_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, tb, hist[(size_t)i].as<int64_t>(), offs[(size_t)i].as<int64_t>(), nh + 1, ctx->stream));
    }
    GSQL_CUDA(ctx, cudaGetLastError());
    // ---- 2. counts[slab][dst] to the host, one AllGather of S x R counts
    std::vector<int64_t> starts((size_t)S * (R + 1), 0), counts((size_t)S * R, 0);
    for (int i = 0; i < S; i++)
        GSQL_CUDA(ctx, cudaMemcpy2DAsync(&starts[(size_t)i * (R + 1)], 8, offs[(size_t)i].p, (size_t)P[(size_t)i].nblocks * 8, 8, (size_t)R,
                                         cudaMemcpyDeviceToHost, ctx->stream));
    GSQL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < S; i++) {
        starts[(size_t)i * (R + 1) + R] = rows_of[(size_t)i];
        for (int d = 0; d < R; d++) counts[(size_t)i * R + d] = starts[(size_t)i * (R + 1) + d + 1] - starts[(size_t)i * (R + 1) + d];
    }
    DevBuf d_counts, d_matrix;
    GSQL_TRY(d_counts.alloc(ctx, (size_t)S * R * 8));
    GSQL_TRY(d_matrix.alloc(ctx, (size_t)R * S * R * 8));
    GSQL_CUDA(ctx, cudaMemcpyAsync(d_counts.p, counts.data(), (size_t)S * R * 8, cudaMemcpyHostToDevice, ctx->stream));
    GSQL_NCCL(ctx, api->AllGather(d_counts.p, d_matrix.p, (size_t)S * R, ncclInt64, comm, ctx->stream));
    std::vector<int64_t> M((size_t)R * S * R);  // M[src][slab][dst]
    GSQL_CUDA(ctx, cudaMemcpyAsync(M.data(), d_matrix.p, M.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    GSQL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int me = ctx->rank;
    auto m_at = [&](int src, int slab, int dst) -> int64_t { return M[((size_t)src * S + slab) * R + dst]; };
    std::vector<int64_t> rbase((size_t)R, 0);  // first receive row of each source
    int64_t total = 0;
    for (int src = 0; src < R; src++) {
        rbase[(size_t)src] = total;
        int64_t from = 0;
        for (int i = 0; i < S; i++) from += m_at(src, i, me);
        if (recv_counts) recv_counts[src] = from;
        total += from;
    }
    *out_rows = total;
    if (total > out_capacity) return gsql_set_error(ctx, GSQL_E_CAPACITY, "all_to_all needs %lld rows, capacity %lld", (long long)total, (long long)out_capacity);
    // ---- 3. scatter slab i (context stream), then its segments leave on the stripe streams while slab i+1 is scattered
    KernelScope ks(ctx, "xchg_alltoall");
    std::vector<int64_t> rdone((size_t)R, 0);  // rows already received from each source (earlier slabs)
    for (int i = 0; i < S; i++) {
        const int64_t lo = (int64_t)i * SR;
        if (rows_of[(size_t)i] > 0) {
            XOut O;
            memset(&O, 0, sizeof(O));
            for (int c = 0; c < s.n_cols; c++) {
                O.data[c] = (char *)sdata[c].p + (size_t)lo * gsql_type_width(s.types[c]);
                if (out->cols[c].nulls) O.nulls[c] = snull[c].as<uint8_t>() + lo;
            }
            k_xchg_scatter<<<P[(size_t)i].nblocks, XBLOCK, s.nparts * sizeof(unsigned long long), ctx->stream>>>(P[(size_t)i], offs[(size_t)i].as<int64_t>(), O);
            GSQL_CUDA(ctx, cudaGetLastError());
        }
        cudaEvent_t ready;
        GSQL_CUDA(ctx, cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        GSQL_CUDA(ctx, cudaEventRecord(ready, ctx->stream));
        for (int k = 0; k < K; k++) {
            ncclComm_t cm = (ncclComm_t)ctx->nccl_extra[k];
            cudaStream_t st = ctx->xstreams[k];
            GSQL_CUDA(ctx, cudaStreamWaitEvent(st, ready, 0));
            GSQL_NCCL(ctx, api->GroupStart());
            for (int c = 0; c < s.n_cols; c++) {
                const size_t w = (size_t)gsql_type_width(s.types[c]);
                for (int peer = 0; peer < R; peer++) {
                    const int64_t ns = counts[(size_t)i * R + peer], nr = m_at(peer, i, me);
                    const int64_t soff = lo + starts[(size_t)i * (R + 1) + peer], roff = rbase[(size_t)peer] + rdone[(size_t)peer];
                    const int64_t s0 = ns * k / K, s1 = ns * (k + 1) / K, r0 = nr * k / K, r1 = nr * (k + 1) / K;
                    if (s1 > s0) GSQL_NCCL(ctx, api->Send((char *)sdata[c].p + (size_t)(soff + s0) * w, (size_t)(s1 - s0) * w, ncclInt8, peer, cm, st));
                    if (r1 > r0) GSQL_NCCL(ctx, api->Recv((char *)out->cols[c].data + (size_t)(roff + r0) * w, (size_t)(r1 - r0) * w, ncclInt8, peer, cm, st));
                    if (out->cols[c].nulls) {
                        if (s1 > s0) GSQL_NCCL(ctx, api->Send((char *)snull[c].p + soff + s0, (size_t)(s1 - s0), ncclInt8, peer, cm, st));
                        if (r1 > r0) GSQL_NCCL(ctx, api->Recv((char *)out->cols[c].nulls + roff + r0, (size_t)(r1 - r0), ncclInt8, peer, cm, st));
                    }
                }
            }
            GSQL_NCCL(ctx, api->GroupEnd());
        }
        GSQL_CUDA(ctx, cudaEventDestroy(ready));
        for (int peer = 0; peer < R; peer++) rdone[(size_t)peer] += m_at(peer, i, me);
    }
    for (int k = 0; k < K; k++) {  // join the stripes back into the context stream
        cudaEvent_t done;
        GSQL_CUDA(ctx, cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        GSQL_CUDA(ctx, cudaEventRecord(done, ctx->xstreams[k]));
        GSQL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, done, 0));
        GSQL_CUDA(ctx, cudaEventDestroy(done));
    }
    out->rows = total;
    return GSQL_OK;
}

extern "C" gsql_status gsql_xchg_all_to_all(gsql_xchg *x, const gsql_batch *in, gsql_batch *out, int64_t out_capacity,
                                            int64_t *out_rows, int64_t *recv_counts) {
    if (!x || !in || !out || !out_rows) return GSQL_E_INVALID;
    gsql_ctx *ctx = x->ctx;
    if (ctx->sticky) return GSQL_E_CUDA;
    const gsql_xchg_spec &s = x->spec;
    NcclApi *api = nccl_api();
    if (!ctx->nccl_comm || !api->ok) return gsql_set_error(ctx, GSQL_E_STATE, "gsql_comm_init has not been called");
    if (s.nparts != ctx->nranks) return gsql_set_error(ctx, GSQL_E_INVALID, "exchange has %d partitions but the communicator has %d ranks", s.nparts, ctx->nranks);
    GSQL_TRY(validate_batch(ctx, in, s.n_cols, s.types));
    GSQL_TRY(validate_batch(ctx, out, s.n_cols, s.types));
    if (in->mem != GSQL_MEM_DEVICE || out->mem != GSQL_MEM_DEVICE) return gsql_set_error(ctx, GSQL_E_INVALID, "all_to_all works on device-resident batches");
    GSQL_CUDA(ctx, cudaSetDevice(ctx->device));
    const int R = ctx->nranks;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    // nullable columns must be nullable on every rank (collective shape): decided by the output buffers
    // ---- 1. partition into contiguous per-destination segments
    StagedBatch sb;
    GSQL_TRY(stage_batch(ctx, in, &sb));
    {
        const int slabs = getenv("GSQL_XCHG_SLABS") ? atoi(getenv("GSQL_XCHG_SLABS")) : 1;  // every rank must use the same value
        if (slabs > 1 && slabs <= 64 && ctx->n_extra > 0) return all_to_all_slabbed(x, sb, out, out_capacity, out_rows, recv_counts, slabs);
    }
    XOut O;
    memset(&O, 0, sizeof(O));
    DevBuf *sdata = x->sdata, *snull = x->snull;
    int64_t srows = in->rows > 0 ? in->rows : 1;
    for (int c = 0; c < s.n_cols; c++) {
        GSQL_TRY(sdata[c].grow(ctx, (size_t)srows * gsql_type_width(s.types[c]), 0));
        O.data[c] = sdata[c].p;
        if (out->cols[c].nulls) {
            GSQL_TRY(snull[c].grow(ctx, (size_t)srows, 0));
            O.nulls[c] = snull[c].as<uint8_t>();
        }
    }
    std::vector<int64_t> send_counts((size_t)R, 0);
    DevBuf offs;
    if (in->rows > 0) GSQL_TRY(partition_device(x, sb, O, &offs, send_counts.data()));
    // ---- 2. exchange the R x R count matrix
    DevBuf d_counts, d_matrix;
    GSQL_TRY(d_counts.alloc(ctx, (size_t)R * 8));
    GSQL_TRY(d_matrix.alloc(ctx, (size_t)R * R * 8));
    GSQL_CUDA(ctx, cudaMemcpyAsync(d_counts.p, send_counts.data(), (size_t)R * 8, cudaMemcpyHostToDevice, ctx->stream));
    GSQL_NCCL(ctx, api->AllGather(d_counts.p, d_matrix.p, (size_t)R, ncclInt64, comm, ctx->stream));
    std::vector<int64_t> matrix((size_t)R * R);
    GSQL_CUDA(ctx, cudaMemcpyAsync(matrix.data(), d_matrix.p, (size_t)R * R * 8, cudaMemcpyDeviceToHost, ctx->stream));
    GSQL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t total = 0;
    std::vector<int64_t> roff((size_t)R), soff((size_t)R);
    for (int src = 0; src < R; src++) {
        roff[(size_t)src] = total;
        int64_t n = matrix[(size_t)src * R + ctx->rank];
        if (recv_counts) recv_counts[src] = n;
        total += n;
    }
    int64_t acc = 0;
    for (int dst = 0; dst < R; dst++) {
        soff[(size_t)dst] = acc;
        acc += send_counts[(size_t)dst];
    }
    *out_rows = total;
    // capacity must be judged identically by all ranks or the collective below would hang: every rank can see the
    // whole matrix, so all of them learn whether anyone overflows... each rank has its own capacity, so the caller
    // contract is: size `out` for the worst case (sum over sources) or retry collectively.
    if (total > out_capacity) return gsql_set_error(ctx, GSQL_E_CAPACITY, "all_to_all needs %lld rows, capacity %lld", (long long)total, (long long)out_capacity);
    // ---- 3. AllToAllv: grouped send/recv per column (and per null mask), striped over 1 + n_extra communicators /
    //         streams so that several NCCL p2p kernels run side by side
    {
        KernelScope ks(ctx, "xchg_alltoall");
        const int K = 1 + ctx->n_extra;
        cudaEvent_t ready;
        GSQL_CUDA(ctx, cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        GSQL_CUDA(ctx, cudaEventRecord(ready, ctx->stream));  // the partitioned send buffers are complete
        for (int k = 0; k < K; k++) {
            ncclComm_t cm = k == 0 ? comm : (ncclComm_t)ctx->nccl_extra[k - 1];
            cudaStream_t st = k == 0 ? ctx->stream : ctx->xstreams[k - 1];
            if (k > 0) GSQL_CUDA(ctx, cudaStreamWaitEvent(st, ready, 0));
            GSQL_NCCL(ctx, api->GroupStart());
            for (int c = 0; c < s.n_cols; c++) {
                size_t w = (size_t)gsql_type_width(s.types[c]);
                for (int peer = 0; peer < R; peer++) {
                    int64_t ns = send_counts[(size_t)peer], nr = matrix[(size_t)peer * R + ctx->rank];
                    // stripe k carries rows [n*k/K, n*(k+1)/K) of every segment
                    int64_t s0 = ns * k / K, s1 = ns * (k + 1) / K, r0 = nr * k / K, r1 = nr * (k + 1) / K;
                    if (s1 > s0) GSQL_NCCL(ctx, api->Send((char *)sdata[c].p + (size_t)(soff[(size_t)peer] + s0) * w, (size_t)(s1 - s0) * w, ncclInt8, peer, cm, st));
                    if (r1 > r0) GSQL_NCCL(ctx, api->Recv((char *)out->cols[c].data + (size_t)(roff[(size_t)peer] + r0) * w, (size_t)(r1 - r0) * w, ncclInt8, peer, cm, st));
                    if (out->cols[c].nulls) {
                        if (s1 > s0) GSQL_NCCL(ctx, api->Send((char *)snull[c].p + soff[(size_t)peer] + s0, (size_t)(s1 - s0), ncclInt8, peer, cm, st));
                        if (r1 > r0) GSQL_NCCL(ctx, api->Recv((char *)out->cols[c].nulls + roff[(size_t)peer] + r0, (size_t)(r1 - r0), ncclInt8, peer, cm, st));
                    }
                }
            }
            GSQL_NCCL(ctx, api->GroupEnd());
        }
        for (int k = 1; k < K; k++) {  // join the stripes back into the context stream
            cudaEvent_t done;
            GSQL_CUDA(ctx, cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
            GSQL_CUDA(ctx, cudaEventRecord(done, ctx->xstreams[k - 1]));
            GSQL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, done, 0));
            GSQL_CUDA(ctx, cudaEventDestroy(done));
        }
        GSQL_CUDA(ctx, cudaEventDestroy(ready));
    }
    out->rows = total;
    return GSQL_OK;
}
