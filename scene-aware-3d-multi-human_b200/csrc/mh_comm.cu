// Library-owned communicator (`mh_set_comm`) and the fused cycle entry points: with it one C call enqueues a
// whole optimisation cycle -- halo exchange, gradients, all-reduce of the shared leaves, update -- on the caller's stream, and no
// Python runs between the kernels of a cycle.
//
// The reference is single-process (mhmocap/predict.py:267-271); the exchanges are this build's frame sharding (sharding.py):
//   * one halo frame (theta, T of the boundary frame: N x 75 floats) with each neighbouring rank, ncclSend / ncclRecv in one group;
//   * ncclAllReduce(sum) of [g_betas | g_xscale | 16 losses] in place in the gradient buffer.
// NCCL is bound at RUN TIME (dlopen "libnccl.so.2" + dlsym): libmhopt.so keeps no link-time dependency on it, and in a process
// that already holds torch's NCCL the same library instance is used.  Without NCCL (or with world == 1) the entry points still
// work: the exchanges are simply skipped / left to the caller (optimizer.py falls back to torch.distributed).
#include "mh_ctx.h"

#include <dlfcn.h>
#include <nccl.h>          // types and enum values only; every function is resolved with dlsym

// The communicator outlives the contexts (ncclCommInitRank takes seconds on an 8-GPU node: a process that fits many sequences creates
// it once, mh_comm_create, and attaches it to every context, mh_set_comm); a context only keeps its neighbours.
struct MhComm {
    void* lib;
    ncclComm_t comm;
    int rank, world, device;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
};

static MhComm* load_nccl(char* err, size_t errlen) {
    MhComm* m = new MhComm();
    memset(m, 0, sizeof(*m));
    m->lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!m->lib) m->lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!m->lib) { snprintf(err, errlen, "NCCL is not available: %s", dlerror()); delete m; return nullptr; }
#define MH_SYM(field, name) do { *(void**)(&m->field) = dlsym(m->lib, name); \
        if (!m->field) { snprintf(err, errlen, "NCCL symbol %s is missing", name); dlclose(m->lib); delete m; return nullptr; } } while (0)
    MH_SYM(GetUniqueId, "ncclGetUniqueId"); MH_SYM(CommInitRank, "ncclCommInitRank"); MH_SYM(CommDestroy, "ncclCommDestroy");
    MH_SYM(AllReduce, "ncclAllReduce"); MH_SYM(Send, "ncclSend"); MH_SYM(Recv, "ncclRecv");
    MH_SYM(GroupStart, "ncclGroupStart"); MH_SYM(GroupEnd, "ncclGroupEnd"); MH_SYM(GetErrorString, "ncclGetErrorString");
#undef MH_SYM
    return m;
}

#define MH_NCCL(ctx, m, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
    snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, (m)->GetErrorString(r_)); return MH_E_CUDA; } } while (0)

struct MhCommLink { MhComm* m; int prev, next; };      // prev / next: ranks owning the adjacent frame ranges, -1 at the ends
static MhCommLink* link_of(mh_ctx* c) { return reinterpret_cast<MhCommLink*>(c->comm); }

// 128-byte NCCL unique id for a new communicator (called on ONE rank; the caller distributes it to the others)
extern "C" int mh_comm_unique_id(mh_ctx* c, uint8_t* out128) {
    if (!c || !out128) return MH_E_ARG;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    MhComm* m = load_nccl(c->err, sizeof(c->err));
    if (!m) return MH_E_STATE;
    ncclUniqueId id;
    ncclResult_t r = m->GetUniqueId(&id);
    if (r != ncclSuccess) { snprintf(c->err, sizeof(c->err), "ncclGetUniqueId: %s", m->GetErrorString(r)); delete m; return MH_E_CUDA; }
    memcpy(out128, &id, 128);
    delete m;                               // the library stays loaded (no dlclose): the id belongs to it
    return MH_OK;
}

// Collective over the mh_dims.world ranks of `c` (its rank / world / device are used; errors are reported through it): creates a
// communicator that is NOT tied to the context.  *handle stays valid until mh_comm_destroy.
extern "C" int mh_comm_create(mh_ctx* c, const uint8_t* unique_id128, void** handle) {
    if (!c || !unique_id128 || !handle) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    const mh_dims& d = c->d;
    if (d.world < 2) MH_FAIL(c, MH_E_ARG, "mh_comm_create: the context was created with world = %d", d.world);
    MhComm* m = load_nccl(c->err, sizeof(c->err));
    if (!m) return MH_E_STATE;
    ncclUniqueId id;
    memcpy(&id, unique_id128, 128);
    ncclResult_t r = m->CommInitRank(&m->comm, d.world, id, d.rank);
    if (r != ncclSuccess) { snprintf(c->err, sizeof(c->err), "ncclCommInitRank: %s", m->GetErrorString(r)); delete m; return MH_E_CUDA; }
    m->rank = d.rank; m->world = d.world; m->device = d.device;
    *handle = m;
    return MH_OK;
}

extern "C" void mh_comm_destroy(void* handle) {
    MhComm* m = reinterpret_cast<MhComm*>(handle);
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->comm) m->CommDestroy(m->comm);
    delete m;
}

// Attach a communicator of mh_comm_create to the context (same rank / world / device).  prev / next = ranks owning the frames before /
// after this rank's range (-1: none).  The context does not own the communicator.
extern "C" int mh_set_comm(mh_ctx* c, void* handle, int32_t prev_rank, int32_t next_rank) {
    if (!c || !handle) return MH_E_ARG;
    const mh_dims& d = c->d;
    MhComm* m = reinterpret_cast<MhComm*>(handle);
    if (m->rank != d.rank || m->world != d.world || m->device != d.device)
        MH_FAIL(c, MH_E_ARG, "mh_set_comm: the communicator is rank %d / %d on device %d, the context rank %d / %d on device %d", m->rank, m->world,
                m->device, d.rank, d.world, d.device);
    if (prev_rank >= d.world || next_rank >= d.world || prev_rank == d.rank || next_rank == d.rank)
        MH_FAIL(c, MH_E_ARG, "mh_set_comm: neighbours (%d, %d) of rank %d in a world of %d", prev_rank, next_rank, d.rank, d.world);
    mh_comm_free(c);
    MhCommLink* l = new MhCommLink();
    l->m = m; l->prev = prev_rank; l->next = next_rank;
    c->comm = l;
    return MH_OK;
}

void mh_comm_free(mh_ctx* c) {
    if (!c->comm) return;
    delete link_of(c);
    c->comm = nullptr;
}

extern "C" int mh_has_comm(mh_ctx* c) { return (c && c->comm) ? 1 : 0; }

// boundary frames of the CURRENT parameters -> the neighbours' halo slots (one grouped send / recv pair per neighbour)
static int comm_halo(mh_ctx* c, cudaStream_t st) {
    MhCommLink* l = link_of(c);
    MhComm* m = l->m;
    const mh_dims& d = c->d;
    const size_t n = (size_t)d.N * MH_HALO;
    MH_TRY(mh_halo_pack(c, st));
    if (l->prev < 0 && l->next < 0) return MH_OK;
    MH_NCCL(c, m, m->GroupStart());
    if (l->prev >= 0) {
        MH_NCCL(c, m, m->Send(c->halo_send, n, ncclFloat32, l->prev, m->comm, st));                 // this rank's FIRST frame
        MH_NCCL(c, m, m->Recv(c->halo_recv, n, ncclFloat32, l->prev, m->comm, st));                 // prev's last frame
    }
    if (l->next >= 0) {
        MH_NCCL(c, m, m->Send(c->halo_send + n, n, ncclFloat32, l->next, m->comm, st));             // this rank's LAST frame
        MH_NCCL(c, m, m->Recv(c->halo_recv + n, n, ncclFloat32, l->next, m->comm, st));             // next's first frame
    }
    MH_NCCL(c, m, m->GroupEnd());
    return MH_OK;
}

static int comm_allreduce_shared(mh_ctx* c, cudaStream_t st) {
    MhComm* m = link_of(c)->m;
    float* shared = c->grads + c->off[MH_P_BETAS];
    MH_NCCL(c, m, m->AllReduce(shared, shared, (size_t)c->d.N * 11 + MH_L_COUNT, ncclFloat32, ncclSum, m->comm, st));
    return MH_OK;
}

// One fit() cycle (optimizer.py:375-587), everything enqueued on `stream`: [halo exchange] -> gradients -> [all-reduce of the shared
// leaves and losses] -> RMSprop step with learning rate `lr`.  Needs mh_set_comm when the context was created with world > 1.
extern "C" int mh_fit_cycle_grads(mh_ctx* c, void* stream) {
    if (!c) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    cudaStream_t st = (cudaStream_t)stream;
    MhCommLink* l = link_of(c);
    if (c->d.world > 1 && !l) MH_FAIL(c, MH_E_STATE, "mh_fit_cycle: a sharded context needs mh_set_comm first");
    if (l) MH_TRY(comm_halo(c, st));
    MH_TRY(mh_fit_grads(c, l ? (l->prev >= 0) : 0, l ? (l->next >= 0) : 0, stream));
    if (l) MH_TRY(comm_allreduce_shared(c, st));
    return MH_OK;
}

extern "C" int mh_fit_cycle(mh_ctx* c, float lr, void* stream) {
    MH_TRY(mh_fit_cycle_grads(c, stream));
    return mh_fit_update(c, lr, stream);
}

// One iteration of the translation init (optimizer.py:743-761): [halo] -> gradients -> [all-reduce] -> Adam step `step` (1-based)
extern "C" int mh_init_cycle(mh_ctx* c, float lr, int32_t step, void* stream) {
    if (!c) return MH_E_ARG;
    cudaSetDevice(c->d.device);
    cudaStream_t st = (cudaStream_t)stream;
    MhCommLink* l = link_of(c);
    if (c->d.world > 1 && !l) MH_FAIL(c, MH_E_STATE, "mh_init_cycle: a sharded context needs mh_set_comm first");
    if (l) MH_TRY(comm_halo(c, st));
    MH_TRY(mh_init_grads(c, l ? (l->prev >= 0) : 0, l ? (l->next >= 0) : 0, stream));
    if (l) MH_TRY(comm_allreduce_shared(c, st));
    return mh_init_update(c, lr, step, stream);
}
