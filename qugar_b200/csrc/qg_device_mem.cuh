// qugar_b200 -- error plumbing, per-device stream + caching allocator, device buffers
// (part of the translation unit qg_engine.cu; included there, in this order)
#pragma once

namespace qg {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int fail(int code, const std::string &msg)
{
  g_last_error = msg;
  return code;
}
#define QG_CUDA(call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) {                                                                           \
      throw EngineError(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? QG_ERR_NO_DEVICE \
                                                                                    : QG_ERR_CUDA,    \
        std::string(#call) + ": " + cudaGetErrorString(e_));                                           \
    }                                                                                                  \
  } while (0)

struct EngineError
{
  int code;
  std::string msg;
  EngineError(int c, const std::string &m) : code(c), msg(m) {}
};

// ------------------------------------------------------------------------------------------------
// per-device context: one stream and a caching allocator.  Work buffers of a rule run to tens of GB and a
// steady workload asks for the same sizes again and again; freed blocks are kept per device and handed
// back to the next request of (nearly) the same size, so that steady-state calls make no driver
// allocation at all and the heap does not fragment.  All work of a device runs on its single stream,
// which orders reuse after the previous user.
// ------------------------------------------------------------------------------------------------
struct DeviceCtx
{
  cudaStream_t stream = nullptr;
  bool ready = false;
  std::mutex m;
  std::multimap<size_t, void *> free_blocks;   // size -> block
  std::unordered_map<void *, size_t> live;     // block -> size
  size_t cached_bytes = 0;
};
static DeviceCtx g_ctx[64];
static std::mutex g_ctx_mtx;

static DeviceCtx &device_ctx(int device)
{
  std::lock_guard<std::mutex> lock(g_ctx_mtx);
  DeviceCtx &c = g_ctx[device];
  if (!c.ready) {
    QG_CUDA(cudaSetDevice(device));
    QG_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.ready = true;
  }
  return c;
}
static cudaStream_t device_stream(int device) { return device_ctx(device).stream; }

static void cache_release_all(DeviceCtx &c)
{
  for (auto &kv : c.free_blocks) cudaFree(kv.second);
  c.free_blocks.clear();
  c.cached_bytes = 0;
}

static void *cache_alloc(int device, size_t bytes)
{
  DeviceCtx &c = device_ctx(device);
  std::lock_guard<std::mutex> lock(c.m);
  bytes = (bytes + 511) & ~(size_t)511;
  // best fit among cached blocks, never wasting more than 25 % (+64 KB) of the block
  auto it = c.free_blocks.lower_bound(bytes);
  if (it != c.free_blocks.end() && it->first <= bytes + bytes / 4 + 65536) {
    void *p = it->second;
    c.live[p] = it->first;
    c.cached_bytes -= it->first;
    c.free_blocks.erase(it);
    return p;
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    QG_CUDA(cudaStreamSynchronize(c.stream));
    cache_release_all(c);
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) throw EngineError(QG_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  c.live[p] = bytes;
  return p;
}
static void cache_free(int device, void *p)
{
  if (!p) return;
  DeviceCtx &c = g_ctx[device];
  std::lock_guard<std::mutex> lock(c.m);
  auto it = c.live.find(p);
  if (it == c.live.end()) return;
  c.free_blocks.emplace(it->second, p);
  c.cached_bytes += it->second;
  c.live.erase(it);
}

// device used by DevBuf allocations made on this thread (set by every API entry point)
static thread_local int g_alloc_device = 0;
static thread_local cudaStream_t g_alloc_stream = nullptr;

// ------------------------------------------------------------------------------------------------
// device buffers
// ------------------------------------------------------------------------------------------------
template<typename T> struct DevBuf
{
  T *p = nullptr;
  size_t n = 0;
  int dev = 0;
  DevBuf() {}
  explicit DevBuf(size_t n_) { alloc(n_); }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), dev(o.dev)
  {
    o.p = nullptr;
    o.n = 0;
  }
  DevBuf &operator=(DevBuf &&o) noexcept
  {
    if (this != &o) {
      release();
      p = o.p;
      n = o.n;
      dev = o.dev;
      o.p = nullptr;
      o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release()
  {
    if (p) cache_free(dev, p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t n_)
  {
    release();
    n = n_;
    dev = g_alloc_device;
    if (n) p = (T *)cache_alloc(dev, n * sizeof(T));
  }
  void zero(cudaStream_t st)
  {
    if (n) QG_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st));
  }
  void upload(const T *h, size_t cnt, cudaStream_t st)
  {
    if (cnt) QG_CUDA(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, st));
  }
  void download(T *h, size_t cnt, cudaStream_t st) const
  {
    if (cnt) QG_CUDA(cudaMemcpyAsync(h, p, cnt * sizeof(T), cudaMemcpyDeviceToHost, st));
  }
};

}  // namespace qg
