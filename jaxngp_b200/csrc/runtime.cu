// runtime.cu -- status channel and per-stream scratch pool of libngp_b200.
#include <dlfcn.h>

#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace ngp {

namespace {
thread_local int t_status = NGP_OK;
thread_local char t_message[512] = {0};

struct Block {
    void *ptr = nullptr;
    size_t bytes = 0;
};
struct Key {
    int device;
    cudaStream_t stream;
    bool operator==(const Key &o) const { return device == o.device && stream == o.stream; }
};
struct KeyHash {
    size_t operator()(const Key &k) const {
        return std::hash<void *>()((void *)k.stream) ^ (std::hash<int>()(k.device) << 1);
    }
};
std::mutex g_mutex;
std::unordered_map<Key, Block, KeyHash> g_blocks;
std::vector<void *> g_retired;  // outgrown blocks, kept alive for captured graphs (see workspace())
}  // namespace

void set_error(int status, const char *fmt, ...) {
    t_status = status;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_message, sizeof(t_message), fmt, ap);
    va_end(ap);
    fprintf(stderr, "[ngp_b200] error %d: %s\n", status, t_message);
}

void clear_error() {
    t_status = NGP_OK;
    t_message[0] = 0;
}

void *workspace(cudaStream_t stream, size_t bytes) {
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(g_mutex);
    Block &b = g_blocks[Key{device, stream}];
    if (b.bytes < bytes) {
        // The old block may still be referenced: by work already enqueued on `stream`, and -- for good -- by CUDA
        // graphs that captured an op which used it (the trainer and the inference renderer replay such graphs for
        // their whole lifetime).  It is therefore retired, not freed: growth happens a handful of times per process
        // (sizes are geometric) and the blocks are a few MB.
        if (b.ptr) g_retired.push_back(b.ptr);
        size_t want = bytes < (1u << 20) ? (1u << 20) : bytes + bytes / 2;
        void *p = nullptr;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            b.ptr = nullptr;
            b.bytes = 0;
            set_error(NGP_ERR_WORKSPACE, "workspace: cudaMalloc(%zu) failed", want);
            return nullptr;
        }
        b.ptr = p;
        b.bytes = want;
    }
    return b.ptr;
}

}  // namespace ngp

// ---- status-returning custom-call form.  XLA's API_VERSION_STATUS_RETURNING (what jax's `custom_call` lowering asks for
// by default) passes an XlaCustomCallStatus* as a fifth argument; a callee with the four-argument signature the
// reference registers (ffi.cc:17-51) simply never looks at it, which is why a failure there can only throw through C
// frames.  The `_status` entry points run the same op and, when it failed and a status object was given, mark it failed
// with the recorded message through XLA's own `XlaCustomCallStatusSetFailure` -- looked up in the process at first use
// (jaxlib exports it from its xla_extension), or handed in with ngp_b200_set_status_failure_fn.
namespace ngp {
namespace {
std::atomic<ngp_set_failure_fn> g_set_failure{nullptr};
std::atomic<bool> g_looked_up{false};
}  // namespace
void report_status(XlaCustomCallStatus *status) {
    if (t_status == NGP_OK || status == nullptr) return;
    ngp_set_failure_fn fn = g_set_failure.load();
    if (!fn && !g_looked_up.exchange(true)) {
        fn = reinterpret_cast<ngp_set_failure_fn>(dlsym(RTLD_DEFAULT, "XlaCustomCallStatusSetFailure"));
        if (fn) g_set_failure.store(fn);
    }
    if (fn) fn(status, t_message, strlen(t_message));
}
}  // namespace ngp

#define NGP_STATUS_FORM(op)                                                                                          \
    extern "C" void op##_status(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len,          \
                                XlaCustomCallStatus *status) {                                                       \
        op(stream, buffers, opaque, opaque_len);                                                                     \
        ngp::report_status(status);                                                                                  \
    }
NGP_STATUS_FORM(ngp_pack_density_into_bits)
NGP_STATUS_FORM(ngp_march_rays)
NGP_STATUS_FORM(ngp_march_rays_inference)
NGP_STATUS_FORM(ngp_morton3d)
NGP_STATUS_FORM(ngp_morton3d_invert)
NGP_STATUS_FORM(ngp_integrate_rays)
NGP_STATUS_FORM(ngp_integrate_rays_backward)
NGP_STATUS_FORM(ngp_integrate_rays_inference)
NGP_STATUS_FORM(ngp_hashgrid_encode)
NGP_STATUS_FORM(ngp_hashgrid_encode_backward)
#undef NGP_STATUS_FORM

extern "C" {
void ngp_b200_set_status_failure_fn(ngp_set_failure_fn fn) { ngp::g_set_failure.store(fn); }
int ngp_b200_abi_version(void) { return NGP_B200_ABI_VERSION; }
int ngp_b200_last_status(void) { return ngp::t_status; }
const char *ngp_b200_last_error(void) { return ngp::t_message; }
void ngp_b200_clear_error(void) { ngp::clear_error(); }
}
