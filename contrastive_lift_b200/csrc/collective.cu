// clift_allreduce_grads: the one collective of the path - a sum-all-reduce of the flat fp32 gradient arena over
// NCCL (NVLink 5 / NVSwitch) on the caller's communicator and stream.  What the reference gets from Lightning's DDPStrategy
// after every manual_backward (trainer/__init__.py:95-108, trainer/train_panopli_tensorf.py:198,220).
//
// The library does not link NCCL: ncclAllReduce is resolved at first use from the libnccl that is already loaded in the
// process (PyTorch's bundled one, or whatever the host application linked), so the same .so works in a process without NCCL as
// long as this entry is never called.
#include <dlfcn.h>

#include "launchers.h"

namespace clift {
namespace {

typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);

struct Nccl {
    nccl_allreduce_fn all_reduce = nullptr;
    nccl_errstr_fn err = nullptr;
    bool tried = false;
};

Nccl& nccl() {
    static Nccl n;
    if (!n.tried) {
        n.tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy the process already uses
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW);
        if (h) {
            n.all_reduce = reinterpret_cast<nccl_allreduce_fn>(dlsym(h, "ncclAllReduce"));
            n.err = reinterpret_cast<nccl_errstr_fn>(dlsym(h, "ncclGetErrorString"));
        }
    }
    return n;
}

}  // namespace
}  // namespace clift

using namespace clift;

extern "C" int32_t clift_allreduce_grads(void* nccl_comm, float* arena, int64_t count, void* stream) {
    CLIFT_CHECK_ARG(nccl_comm != nullptr && count >= 0, "null communicator or negative count");
    if (count == 0) return CLIFT_OK;
    CLIFT_CHECK_ARG(arena != nullptr, "null arena");
    Nccl& n = nccl();
    if (!n.all_reduce) {
        set_error("clift_allreduce_grads: no libnccl.so.2 with ncclAllReduce is loadable in this process");
        return CLIFT_ERR_UNSUPPORTED;
    }
    constexpr int kNcclFloat32 = 7, kNcclSum = 0;      // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0)
    const int rc = n.all_reduce(arena, arena, (size_t)count, kNcclFloat32, kNcclSum, nccl_comm, (cudaStream_t)stream);
    if (rc != 0) {
        set_error("clift_allreduce_grads: ncclAllReduce -> %s", n.err ? n.err(rc) : "error");
        return CLIFT_ERR_CUDA;
    }
    count_launch();
    return CLIFT_OK;
}
