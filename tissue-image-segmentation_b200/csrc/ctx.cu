// Context, error reporting and the call-scoped device workspace of libtiseg_b200.so.
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace tiseg {

static thread_local std::string g_err;

void set_error(const std::string& msg) { g_err = msg; }

int fail(const char* where, cudaError_t e) {
    g_err = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return TISEG_ERR_CUDA;
}

void begin_call(tiseg_ctx* c) {
    cudaSetDevice(c->device);
    // coalesce the arena if the previous call spilled into extra blocks
    if (c->blocks.size() > 1) {
        size_t total = 0;
        for (auto& b : c->blocks) total += b.cap;
        cudaStreamSynchronize(c->stream);
        for (auto& b : c->blocks) cudaFree(b.p);
        c->blocks.clear();
        char* p = nullptr;
        if (cudaMalloc(&p, total) == cudaSuccess) c->blocks.push_back({p, total});
        else cudaGetLastError();
    }
    c->cur_block = 0;
    c->cur_off = 0;
    c->pending.clear();
}

void* ws_alloc(tiseg_ctx* c, size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    if (bytes == 0) bytes = 256;
    while (c->cur_block < c->blocks.size()) {
        auto& b = c->blocks[c->cur_block];
        if (c->cur_off + bytes <= b.cap) {
            void* r = b.p + c->cur_off;
            c->cur_off += bytes;
            return r;
        }
        c->cur_block++;
        c->cur_off = 0;
    }
    size_t cap = bytes > (size_t(256) << 20) ? bytes : (size_t(256) << 20);
    char* p = nullptr;
    if (cudaMalloc(&p, cap) != cudaSuccess) {
        cudaGetLastError();
        set_error("workspace cudaMalloc failed");
        return nullptr;
    }
    c->blocks.push_back({p, cap});
    c->cur_block = c->blocks.size() - 1;
    c->cur_off = bytes;
    return p;
}

bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

const void* in_ptr(tiseg_ctx* c, const void* p, size_t bytes) {
    if (!p) return nullptr;
    if (is_device_ptr(p)) return p;
    void* d = ws_alloc(c, bytes);
    if (!d) return nullptr;
    if (cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
        fail("H2D", cudaGetLastError());
        return nullptr;
    }
    return d;
}

void* out_ptr(tiseg_ctx* c, void* p, size_t bytes) {
    if (!p) return nullptr;
    if (is_device_ptr(p)) return p;
    void* d = ws_alloc(c, bytes);
    if (!d) return nullptr;
    c->pending.push_back({p, d, bytes});
    return d;
}

void* inout_ptr(tiseg_ctx* c, void* p, size_t bytes) {
    if (!p) return nullptr;
    if (is_device_ptr(p)) return p;
    void* d = ws_alloc(c, bytes);
    if (!d) return nullptr;
    if (cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
        fail("H2D", cudaGetLastError());
        return nullptr;
    }
    c->pending.push_back({p, d, bytes});
    return d;
}

int end_call(tiseg_ctx* c) {
    if (c->pending.empty()) return TISEG_OK;
    for (auto& q : c->pending)
        TISEG_CHECK(cudaMemcpyAsync(q.host, q.dev, q.bytes, cudaMemcpyDeviceToHost, c->stream));
    c->pending.clear();
    TISEG_CHECK(cudaStreamSynchronize(c->stream));
    return TISEG_OK;
}

int zero(tiseg_ctx* c, void* p, size_t bytes) {
    TISEG_CHECK(cudaMemsetAsync(p, 0, bytes, c->stream));
    return TISEG_OK;
}

}  // namespace tiseg

extern "C" {

int tiseg_create(tiseg_ctx** out, int device) {
    if (!out) { tiseg::set_error("tiseg_create: null out"); return TISEG_ERR_ARG; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        tiseg::set_error("tiseg_create: no CUDA device (libtiseg_b200 has no CPU fallback)");
        return TISEG_ERR_NOGPU;
    }
    if (device < 0 || device >= n) { tiseg::set_error("tiseg_create: bad device index"); return TISEG_ERR_ARG; }
    TISEG_CHECK(cudaSetDevice(device));
    tiseg_ctx* c = new tiseg_ctx();
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return tiseg::fail("cudaStreamCreate", e); }
    c->own_stream = true;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = c;
    return TISEG_OK;
}

int tiseg_destroy(tiseg_ctx* c) {
    if (!c) return TISEG_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& b : c->blocks) cudaFree(b.p);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return TISEG_OK;
}

int tiseg_set_stream(tiseg_ctx* c, void* s) {
    if (!c) { tiseg::set_error("null ctx"); return TISEG_ERR_ARG; }
    cudaSetDevice(c->device);
    if (c->own_stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
        c->own_stream = false;
    }
    c->stream = (cudaStream_t)s;   // NULL == the legacy default stream (what torch uses by default)
    return TISEG_OK;
}

int tiseg_synchronize(tiseg_ctx* c) {
    if (!c) { tiseg::set_error("null ctx"); return TISEG_ERR_ARG; }
    TISEG_CHECK(cudaStreamSynchronize(c->stream));
    return TISEG_OK;
}

const char* tiseg_last_error(void) { return tiseg::g_err.c_str(); }

long long tiseg_launch_count(tiseg_ctx* c) { return c ? c->launches : 0; }

int tiseg_version(void) { return 100; }

}  // extern "C"
