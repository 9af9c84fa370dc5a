// bmo_comm.cu -- the one exchange step of the path behind the C ABI: the Photodetector field of a beamlet bundle
// that was sharded over several GPUs is the sum of the ranks' partial fields (the reference adds the beamlet
// fields serially, `pd.field[i, j] += ...`, Photodetector.jl:103; its `for beam in beams(bg)` loop, System.jl:463-468,
// is what gets sharded).  One complex128 all-reduce (2 n^2 doubles) over NVLink / NVSwitch through NCCL.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that libbmo.so has no link-time dependency on it: a host
// that never shards never needs the library, and inside a process that already carries one (PyTorch's bundled
// NCCL) the same copy is used.  BMO_NCCL_LIB overrides the library path.
//
// Two ways to build the communicator, matching the two host models:
//   * one process per GPU (torchrun, MPI, Julia Distributed): rank 0 calls bmo_comm_unique_id, the 128 bytes travel by
//     whatever transport the host has, every rank calls bmo_comm_init;
//   * one process driving n GPUs (a single Julia session, SURVEY 8(b) `bmo_init(n_gpus)`): bmo_comm_init_local over the
//     contexts of the n devices, bmo_pd_allreduce_local = the grouped all-reduce.
#include <dlfcn.h>
#include <mutex>
#include "bmo_host.cuh"

using namespace bmo;

namespace {
// the slice of nccl.h this file uses (stable since NCCL 2.0; checked against 2.27.3 / 2.28.9)
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*fn_GetUniqueId)(NcclUniqueId*);
typedef int (*fn_CommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_CommInitAll)(NcclComm*, int, const int*);
typedef int (*fn_CommDestroy)(NcclComm);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int /*dtype*/, int /*op*/, NcclComm, cudaStream_t);
typedef int (*fn_Group)(void);
typedef const char* (*fn_ErrStr)(int);
typedef int (*fn_GetVersion)(int*);
constexpr int kNcclDouble = 8, kNcclSum = 0;

struct Nccl {
    void* so = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommInitAll CommInitAll = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_Group GroupStart = nullptr, GroupEnd = nullptr;
    fn_ErrStr GetErrorString = nullptr;
    fn_GetVersion GetVersion = nullptr;
    std::string why;   // load error
};
Nccl g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
    const char* env = getenv("BMO_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        g_nccl.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.so) break;
        g_nccl.why = dlerror();
    }
    if (!g_nccl.so) return;
#define SYM(field, name) g_nccl.field = (decltype(g_nccl.field))dlsym(g_nccl.so, name); if (!g_nccl.field) { g_nccl.why = std::string("symbol missing: ") + name; g_nccl.so = nullptr; return; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommInitAll, "ncclCommInitAll")
    SYM(CommDestroy, "ncclCommDestroy") SYM(AllReduce, "ncclAllReduce") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString") SYM(GetVersion, "ncclGetVersion")
#undef SYM
}
int32_t need_nccl() {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.so) return fail(BMO_ENCCL, "NCCL is not available (dlopen libnccl.so.2: " + g_nccl.why + "); set BMO_NCCL_LIB");
    return BMO_OK;
}
#define BMO_NCCL(call)                                                                                          \
    do {                                                                                                        \
        int r_ = (call);                                                                                        \
        if (r_ != 0) return fail(BMO_ENCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_));             \
    } while (0)
}  // namespace

struct bmo_comm {
    bmo_ctx* ctx = nullptr;
    NcclComm comm = nullptr;
    int32_t rank = 0, n_ranks = 1;
    double* stage = nullptr;     // device staging buffer for host fields
    size_t stage_cap = 0;        // doubles
};

int32_t bmo_comm_unique_id(uint8_t* id) {
    if (!id) return fail(BMO_EINVAL, "bmo_comm_unique_id: NULL");
    int32_t rc = need_nccl();
    if (rc) return rc;
    NcclUniqueId u;
    BMO_NCCL(g_nccl.GetUniqueId(&u));
    std::memcpy(id, u.internal, BMO_COMM_ID_BYTES);
    return BMO_OK;
}

int32_t bmo_comm_init(bmo_ctx* ctx, int32_t n_ranks, int32_t rank, const uint8_t* id, bmo_comm** out) {
    if (!ctx || !id || !out) return fail(BMO_EINVAL, "bmo_comm_init: NULL argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(BMO_EINVAL, "bmo_comm_init: rank outside [0, n_ranks)");
    int32_t rc = need_nccl();
    if (rc) return rc;
    BMO_CUDA(cudaSetDevice(ctx->device));
    NcclUniqueId u;
    std::memcpy(u.internal, id, BMO_COMM_ID_BYTES);
    bmo_comm* c = new bmo_comm();
    c->ctx = ctx; c->rank = rank; c->n_ranks = n_ranks;
    int r = g_nccl.CommInitRank(&c->comm, n_ranks, u, rank);
    if (r != 0) { delete c; return fail(BMO_ENCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); }
    *out = c;
    return BMO_OK;
}

int32_t bmo_comm_init_local(int32_t n, bmo_ctx* const* ctxs, bmo_comm** comms) {
    if (n < 1 || !ctxs || !comms) return fail(BMO_EINVAL, "bmo_comm_init_local: bad arguments");
    int32_t rc = need_nccl();
    if (rc) return rc;
    std::vector<int> devs((size_t)n);
    for (int k = 0; k < n; k++) {
        if (!ctxs[k]) return fail(BMO_EINVAL, "bmo_comm_init_local: NULL context");
        devs[k] = ctxs[k]->device;
        for (int j = 0; j < k; j++) if (devs[j] == devs[k]) return fail(BMO_EINVAL, "bmo_comm_init_local: two contexts share a device");
    }
    std::vector<NcclComm> cs((size_t)n, nullptr);
    BMO_NCCL(g_nccl.CommInitAll(cs.data(), n, devs.data()));
    for (int k = 0; k < n; k++) {
        bmo_comm* c = new bmo_comm();
        c->ctx = ctxs[k]; c->comm = cs[k]; c->rank = k; c->n_ranks = n;
        comms[k] = c;
    }
    return BMO_OK;
}

int32_t bmo_comm_info(bmo_comm* c, int32_t* rank, int32_t* n_ranks, int32_t* nccl_version) {
    if (!c) return fail(BMO_EINVAL, "bmo_comm_info: NULL");
    if (rank) *rank = c->rank;
    if (n_ranks) *n_ranks = c->n_ranks;
    if (nccl_version) { int v = 0; g_nccl.GetVersion(&v); *nccl_version = v; }
    return BMO_OK;
}

// device pointer of the field on this rank: the caller's (BMO_INPUT_DEVICE) or the staging copy of a host field
static int32_t stage_field(bmo_comm* c, double* field, size_t n_dbl, bool on_dev, double** dptr) {
    if (on_dev) { *dptr = field; return BMO_OK; }
    cudaStream_t st = c->ctx->stream;
    if (c->stage_cap < n_dbl) {
        if (c->stage) cudaFree(c->stage);
        c->stage = nullptr; c->stage_cap = 0;
        BMO_CUDA(cudaMalloc((void**)&c->stage, n_dbl * sizeof(double)));
        c->stage_cap = n_dbl;
    }
    BMO_CUDA(cudaMemcpyAsync(c->stage, field, n_dbl * sizeof(double), cudaMemcpyHostToDevice, st));
    *dptr = c->stage;
    return BMO_OK;
}

int32_t bmo_pd_allreduce(bmo_comm* c, double* field, int64_t n_complex, uint32_t flags) {
    if (!c || !field || n_complex <= 0) return fail(BMO_EINVAL, "bmo_pd_allreduce: bad arguments");
    BMO_CUDA(cudaSetDevice(c->ctx->device));
    cudaStream_t st = c->ctx->stream;
    const bool on_dev = flags & BMO_INPUT_DEVICE;
    const size_t n_dbl = (size_t)2 * (size_t)n_complex;
    double* d = nullptr;
    int32_t rc = stage_field(c, field, n_dbl, on_dev, &d);
    if (rc) return rc;
    if (c->n_ranks > 1) BMO_NCCL(g_nccl.AllReduce(d, d, n_dbl, kNcclDouble, kNcclSum, c->comm, st));
    if (!on_dev) BMO_CUDA(cudaMemcpyAsync(field, d, n_dbl * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (!on_dev || (flags & BMO_COMM_SYNC)) BMO_CUDA(cudaStreamSynchronize(st));
    return BMO_OK;
}

int32_t bmo_pd_allreduce_local(int32_t n, bmo_comm* const* comms, double* const* fields, int64_t n_complex, uint32_t flags) {
    if (n < 1 || !comms || !fields || n_complex <= 0) return fail(BMO_EINVAL, "bmo_pd_allreduce_local: bad arguments");
    const bool on_dev = flags & BMO_INPUT_DEVICE;
    const size_t n_dbl = (size_t)2 * (size_t)n_complex;
    std::vector<double*> d((size_t)n, nullptr);
    int32_t rc;
    for (int k = 0; k < n; k++) {
        if (!comms[k] || !fields[k]) return fail(BMO_EINVAL, "bmo_pd_allreduce_local: NULL entry");
        BMO_CUDA(cudaSetDevice(comms[k]->ctx->device));
        if ((rc = stage_field(comms[k], fields[k], n_dbl, on_dev, &d[k]))) return rc;
    }
    if (n > 1) {
        BMO_NCCL(g_nccl.GroupStart());
        for (int k = 0; k < n; k++) {
            int r = g_nccl.AllReduce(d[k], d[k], n_dbl, kNcclDouble, kNcclSum, comms[k]->comm, comms[k]->ctx->stream);
            if (r != 0) { g_nccl.GroupEnd(); return fail(BMO_ENCCL, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r)); }
        }
        BMO_NCCL(g_nccl.GroupEnd());
    }
    for (int k = 0; k < n; k++) {
        BMO_CUDA(cudaSetDevice(comms[k]->ctx->device));
        if (!on_dev) BMO_CUDA(cudaMemcpyAsync(fields[k], d[k], n_dbl * sizeof(double), cudaMemcpyDeviceToHost, comms[k]->ctx->stream));
    }
    for (int k = 0; k < n; k++) {
        BMO_CUDA(cudaSetDevice(comms[k]->ctx->device));
        if (!on_dev || (flags & BMO_COMM_SYNC)) BMO_CUDA(cudaStreamSynchronize(comms[k]->ctx->stream));
    }
    return BMO_OK;
}

int32_t bmo_comm_free(bmo_comm* c) {
    if (!c) return BMO_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->comm && g_nccl.so) g_nccl.CommDestroy(c->comm);
    if (c->stage) cudaFree(c->stage);
    delete c;
    return BMO_OK;
}
