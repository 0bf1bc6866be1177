// runtime.cu -- status channel and per-stream scratch pool of libngp_b200.
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

extern "C" {
int ngp_b200_abi_version(void) { return NGP_B200_ABI_VERSION; }
int ngp_b200_last_status(void) { return ngp::t_status; }
const char *ngp_b200_last_error(void) { return ngp::t_message; }
void ngp_b200_clear_error(void) { ngp::clear_error(); }
}
