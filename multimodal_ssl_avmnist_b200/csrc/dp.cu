// Data-parallel exchange inside the C ABI: one NCCL communicator per process (one process per GPU), used
// for exactly two all-reduces per step -- the flat gradient arena (in 1-3 slices) and the [D] teacher column sums.  The
// reference gets the same exchange from Lightning's strategy="ddp" (run_dino.py:359).  Calls are asynchronous on the given
// stream and can be captured into a CUDA graph together with the kernels around them.
//
// NCCL is bound at run time (dlopen, no link-time dependency): the library already loaded in the process (the one PyTorch
// ships) is preferred so that a process never holds two NCCL versions; a stand-alone C client gets the system libnccl.so.2.
// Only NCCL's stable C interface is used (ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclCommDestroy).
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace b200 {
namespace {

typedef struct { char internal[128]; } NcclUniqueId;         // NCCL_UNIQUE_ID_BYTES = 128 (nccl.h)
typedef void* NcclComm;
constexpr int NCCL_FLOAT32 = 7, NCCL_SUM = 0;                // ncclDataType_t / ncclRedOp_t values (nccl.h)

struct Nccl {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
Nccl g_nccl;
NcclComm g_comm = nullptr;
int g_rank = 0, g_world = 1;

int load_nccl() {
    if (g_nccl.handle) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy already in the process (PyTorch's), if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("dp: cannot load libnccl.so.2 (%s)", dlerror());
        return B200_E_ARG;
    }
    g_nccl.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
    g_nccl.GetVersion = (int (*)(int*))dlsym(h, "ncclGetVersion");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        set_error("dp: libnccl.so.2 lacks a required symbol");
        return B200_E_ARG;
    }
    g_nccl.handle = h;
    return 0;
}

int nccl_status(int rc, const char* what) {
    if (rc == 0) return 0;
    set_error("%s: NCCL error %d (%s)", what, rc, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return 1000 + rc;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_dp_nccl_version(void) {
    if (load_nccl()) return -1;
    int v = 0;
    if (g_nccl.GetVersion) g_nccl.GetVersion(&v);
    return v;
}

int b200_dp_unique_id(char* id_out) {
    B200_REQUIRE(id_out, B200_E_ARG, "dp_unique_id: null pointer");
    if (int rc = load_nccl()) return rc;
    NcclUniqueId id;
    if (int rc = nccl_status(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) return rc;
    memcpy(id_out, id.internal, sizeof(id.internal));
    return 0;
}

int b200_dp_init(const char* id, int rank, int world) {
    B200_REQUIRE(id && world >= 1 && rank >= 0 && rank < world, B200_E_ARG, "dp_init: bad arguments");
    B200_REQUIRE(g_comm == nullptr, B200_E_ARG, "dp_init: a communicator already exists (call b200_dp_destroy first)");
    if (int rc = load_nccl()) return rc;
    NcclUniqueId uid;
    memcpy(uid.internal, id, sizeof(uid.internal));
    if (int rc = nccl_status(g_nccl.CommInitRank(&g_comm, world, uid, rank), "ncclCommInitRank")) {
        g_comm = nullptr;
        return rc;
    }
    g_rank = rank;
    g_world = world;
    return 0;
}

int b200_dp_world(void) { return g_comm ? g_world : 1; }
int b200_dp_rank(void) { return g_comm ? g_rank : 0; }

static int allreduce_sum(float* buf, int64_t n, void* stream, const char* what) {
    B200_REQUIRE(buf && n > 0, B200_E_ARG, "%s: bad arguments", what);
    B200_REQUIRE(g_comm != nullptr, B200_E_ARG, "%s: no communicator (call b200_dp_init)", what);
    return nccl_status(g_nccl.AllReduce(buf, buf, (size_t)n, NCCL_FLOAT32, NCCL_SUM, g_comm, as_stream(stream)), what);
}

int b200_dp_allreduce_grads(float* grad, int64_t n, void* stream) { return allreduce_sum(grad, n, stream, "dp_allreduce_grads"); }

int b200_dp_allreduce_center(float* colsum, int D, void* stream) { return allreduce_sum(colsum, D, stream, "dp_allreduce_center"); }

int b200_dp_destroy(void) {
    if (g_comm) {
        int rc = g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
        g_world = 1;
        g_rank = 0;
        return nccl_status(rc, "ncclCommDestroy");
    }
    return 0;
}

}  // extern "C"
