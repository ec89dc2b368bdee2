// fused.cu — NVRTC compile + module cache + launch of lowered fused programs.
//
// Mirrors the role of the wgpu provider's pipeline cache (`get_or_create_pipeline`, keyed by a hash of the
// shader bytes + layout tag: backend/wgpu/provider/ops/elementwise.rs:1608-1626) and its fused dispatch
// (`fused_elementwise_exec` :1567-1848, `fused_reduction` via reduction/autotune.rs:7-89), re-designed for
// CUDA: the lowered source is compiled once by NVRTC for sm_100a (-fmad=false), the cubin is cached in
// memory by key and on disk by content hash, and launches go through the driver entry points that the
// static CUDA runtime resolves (no link-time libcuda dependency, so the library loads on a CPU-only box).
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>

#include "common.h"

namespace rm {

namespace {

struct DriverApi {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                           CUstream, void**, void**) = nullptr;
  CUresult (*LaunchKernelEx)(const CUlaunchConfig*, CUfunction, void**, void**) = nullptr;  // optional (programmatic dependent launch)
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*CtxGetCurrent)(CUcontext*) = nullptr;
  CUresult (*CtxSetCurrent)(CUcontext) = nullptr;
  bool ok = false;
};

DriverApi& driver() {
  static DriverApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
    };
    bool ok = true;
    ok &= get("cuModuleLoadData", (void**)&api.ModuleLoadData);
    ok &= get("cuModuleGetFunction", (void**)&api.ModuleGetFunction);
    ok &= get("cuModuleUnload", (void**)&api.ModuleUnload);
    ok &= get("cuLaunchKernel", (void**)&api.LaunchKernel);
    ok &= get("cuGetErrorString", (void**)&api.GetErrorString);
    ok &= get("cuCtxGetCurrent", (void**)&api.CtxGetCurrent);
    ok &= get("cuCtxSetCurrent", (void**)&api.CtxSetCurrent);
    if (!get("cuLaunchKernelEx", (void**)&api.LaunchKernelEx)) api.LaunchKernelEx = nullptr;
    if (getenv("RUNMAT_B200_NO_PDL")) api.LaunchKernelEx = nullptr;
    cudaGetLastError();
    api.ok = ok;
  });
  return api;
}

const char* cu_err(CUresult r) {
  const char* s = nullptr;
  if (driver().GetErrorString) driver().GetErrorString(r, &s);
  return s ? s : "unknown CUDA driver error";
}

uint64_t fnv1a(const std::string& s) {
  uint64_t h = 1469598103934665603ULL;
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
  return h;
}

std::string cache_dir() {
  if (const char* e = getenv("RUNMAT_B200_KCACHE")) return e;
  Dl_info info;
  if (dladdr((void*)&fnv1a, &info) && info.dli_fname) {
    std::string path(info.dli_fname);
    size_t slash = path.rfind('/');
    if (slash != std::string::npos) return path.substr(0, slash) + "/_kcache";
  }
  return "/tmp/runmat_b200_kcache";
}

const char* kNvrtcOpts[] = {"--gpu-architecture=sm_100a", "-fmad=false", "-lineinfo", "--std=c++17"};
const char* kOptsTag = "sm_100a|-fmad=false|v4";

}  // namespace

struct Kernel {
  CUmodule module = nullptr;
  CUfunction fn = nullptr;
};

struct FusedCache {
  std::mutex mu;
  std::unordered_map<std::string, Kernel> kernels;  // key -> loaded function
  CUcontext ctx = nullptr;
  void* tickets = nullptr;  // zero-initialised u32[kTicketCap]; last block resets its ticket
  static constexpr uint32_t kTicketCap = 65536;
};

rm_status compile_cuda_to_cubin(const std::string& src, const char* name, std::vector<char>* cubin, std::string* log) {
  // disk cache (content-addressed): survives processes; ships to the GPU box with the snapshot
  const std::string dir = cache_dir();
  char hex[32];
  snprintf(hex, sizeof hex, "%016llx", (unsigned long long)fnv1a(src + kOptsTag));
  const std::string path = dir + "/" + hex + ".cubin";
  const bool use_disk = !getenv("RUNMAT_B200_NO_KCACHE");
  if (use_disk) {
    std::ifstream f(path, std::ios::binary);
    if (f) {
      cubin->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
      if (!cubin->empty()) return RM_OK;
    }
  }
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, src.c_str(), name, 0, nullptr, nullptr) != NVRTC_SUCCESS)
    return fail(RM_COMPILE_ERROR, "nvrtcCreateProgram failed");
  nvrtcResult r = nvrtcCompileProgram(prog, (int)(sizeof(kNvrtcOpts) / sizeof(kNvrtcOpts[0])), kNvrtcOpts);
  size_t lsz = 0;
  nvrtcGetProgramLogSize(prog, &lsz);
  std::string l(lsz, '\0');
  if (lsz > 1) nvrtcGetProgramLog(prog, &l[0]);
  if (log) *log = l;
  if (r != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    if (getenv("RUNMAT_DEBUG_FUSION")) fprintf(stderr, "[runmat_b200] NVRTC failure for %s:\n%s\n---- source ----\n%s\n", name, l.c_str(), src.c_str());
    return fail(RM_COMPILE_ERROR, "NVRTC compile of %s failed: %s", name, l.c_str());
  }
  size_t csz = 0;
  nvrtcGetCUBINSize(prog, &csz);
  cubin->resize(csz);
  nvrtcGetCUBIN(prog, cubin->data());
  nvrtcDestroyProgram(&prog);
  if (use_disk && csz) {
    mkdir(dir.c_str(), 0755);
    std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
    std::ofstream f(tmp, std::ios::binary);
    if (f) {
      f.write(cubin->data(), (std::streamsize)cubin->size());
      f.close();
      rename(tmp.c_str(), path.c_str());
    }
  }
  return RM_OK;
}

rm_status fused_cache_create(rm_provider* p) {
  p->fused = new FusedCache();
  DriverApi& d = driver();
  if (!d.ok) return fail(RM_NO_DEVICE, "CUDA driver entry points unavailable");
  RM_CUDA(cudaFree(0));  // binds the primary context to this thread
  if (d.CtxGetCurrent(&p->fused->ctx) != CUDA_SUCCESS || !p->fused->ctx) return fail(RM_NO_DEVICE, "no current CUDA context");
  RM_CUDA(cudaMalloc(&p->fused->tickets, FusedCache::kTicketCap * sizeof(uint32_t)));
  RM_CUDA(cudaMemset(p->fused->tickets, 0, FusedCache::kTicketCap * sizeof(uint32_t)));
  return RM_OK;
}

void fused_cache_destroy(rm_provider* p) {
  if (!p->fused) return;
  DriverApi& d = driver();
  for (auto& kv : p->fused->kernels)
    if (kv.second.module && d.ModuleUnload) d.ModuleUnload(kv.second.module);
  if (p->fused->tickets) cudaFree(p->fused->tickets);
  delete p->fused;
  p->fused = nullptr;
}

namespace {

// Looks up (or lowers+compiles+loads) the kernel for `key`. `make_src` is only invoked on a miss.
template <typename MakeSrc>
rm_status get_kernel(rm_provider* p, const std::string& key, const char* entry, MakeSrc make_src, Kernel* out) {
  FusedCache& c = *p->fused;
  DriverApi& d = driver();
  std::lock_guard<std::mutex> lk(c.mu);
  CUcontext cur = nullptr;
  d.CtxGetCurrent(&cur);
  if (cur != c.ctx) d.CtxSetCurrent(c.ctx);
  auto it = c.kernels.find(key);
  if (it != c.kernels.end()) {
    p->cache_hits.fetch_add(1, std::memory_order_relaxed);
    *out = it->second;
    return RM_OK;
  }
  p->cache_misses.fetch_add(1, std::memory_order_relaxed);
  std::string src = make_src();
  if (getenv("RUNMAT_DEBUG_DUMP_FUSED_CUDA")) fprintf(stderr, "---- fused CUDA (%s) ----\n%s\n", entry, src.c_str());
  std::vector<char> cubin;
  std::string log;
  RM_TRY(compile_cuda_to_cubin(src, entry, &cubin, &log));
  Kernel k;
  CUresult r = d.ModuleLoadData(&k.module, cubin.data());
  if (r != CUDA_SUCCESS) return fail(RM_COMPILE_ERROR, "cuModuleLoadData failed: %s", cu_err(r));
  r = d.ModuleGetFunction(&k.fn, k.module, entry);
  if (r != CUDA_SUCCESS) return fail(RM_COMPILE_ERROR, "cuModuleGetFunction(%s) failed: %s", entry, cu_err(r));
  c.kernels.emplace(key, k);
  *out = k;
  return RM_OK;
}

rm_status launch(rm_provider* p, const Kernel& k, dim3 grid, dim3 block, void** args) {
  DriverApi& d = driver();
  CUcontext cur = nullptr;
  d.CtxGetCurrent(&cur);
  if (cur != p->fused->ctx) d.CtxSetCurrent(p->fused->ctx);
  CUresult r;
  if (d.LaunchKernelEx && p->launch_overlap) {
    // programmatic stream serialisation: the generated kernels open with griddepcontrol.launch_dependents + griddepcontrol.wait
    // (fusion_lower.cpp RM_PDL_PROLOGUE), so this kernel's CTAs may become resident while the previous kernel drains
    CUlaunchAttribute attr;
    memset(&attr, 0, sizeof attr);
    attr.id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
    attr.value.programmaticStreamSerializationAllowed = 1;
    CUlaunchConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDimX = grid.x; cfg.gridDimY = grid.y; cfg.gridDimZ = grid.z;
    cfg.blockDimX = block.x; cfg.blockDimY = block.y; cfg.blockDimZ = block.z;
    cfg.sharedMemBytes = 0;
    cfg.hStream = (CUstream)p->stream;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    r = d.LaunchKernelEx(&cfg, k.fn, args, nullptr);
  } else {
    r = d.LaunchKernel(k.fn, grid.x, grid.y, grid.z, block.x, block.y, block.z, 0, (CUstream)p->stream, args, nullptr);
  }
  if (r != CUDA_SUCCESS) return fail(RM_ERROR, "cuLaunchKernel failed: %s", cu_err(r));
  count_launch(p);
  return RM_OK;
}

}  // namespace

rm_status run_elementwise_program(rm_provider* p, const ElementwiseProgram& prog, const std::string& key,
                                  const rm_handle* inputs, uint32_t n_inputs, const uint64_t* out_shape,
                                  uint32_t rank, uint64_t len, rm_handle* outs) {
  RM_REQUIRE(n_inputs == prog.n_inputs, RM_INVALID_ARG, "fused_elementwise: shader declares %u inputs, got %u", prog.n_inputs, n_inputs);
  RM_REQUIRE(n_inputs <= 24, RM_UNSUPPORTED, "fused_elementwise: too many inputs (%u)", n_inputs);
  RM_REQUIRE(rank <= RM_MAX_RANK, RM_UNSUPPORTED, "fused_elementwise: rank %u exceeds %d", rank, RM_MAX_RANK);
  RM_REQUIRE((p->precision == RM_F64) == (prog.scalar_ty == "f64"), RM_INVALID_ARG,
             "fused_elementwise: shader scalar type %s does not match provider precision", prog.scalar_ty.c_str());
  RM_REQUIRE(shape_elems(out_shape, rank) == len, RM_INVALID_ARG, "fused_elementwise: len %llu does not match output shape", (unsigned long long)len);
  RM_REQUIRE(len > 0, RM_ERROR, "fusion: zero-length execution not supported");  // fusion_exec.rs:272-274

  std::vector<void*> in_ptr(n_inputs);
  std::vector<std::vector<uint64_t>> in_shape(n_inputs, std::vector<uint64_t>(rank, 1)), in_stride(n_inputs, std::vector<uint64_t>(rank, 0));
  bool flat = true;
  uint32_t scalar_mask = 0;
  for (uint32_t k = 0; k < n_inputs; ++k) {
    uint64_t elems = 0;
    RM_TRY(resolve(p, &inputs[k], &in_ptr[k], &elems));
    const rm_handle& h = inputs[k];
    RM_REQUIRE(h.rank <= rank || elems == 1, RM_INVALID_ARG, "fused_elementwise: input %u rank %u exceeds output rank %u", k, h.rank, rank);
    // front-pad to the output rank (wgpu provider: elementwise.rs:1671-1677)
    if (elems == 1) {
      // 1-element tensors broadcast against anything (scalars uploaded by the executor)
    } else {
      uint32_t pad = rank - h.rank;
      for (uint32_t d = 0; d < h.rank; ++d) in_shape[k][pad + d] = h.shape[d];
    }
    uint64_t s = 1;
    bool same = true;
    for (uint32_t d = 0; d < rank; ++d) {
      uint64_t dim = in_shape[k][d];
      RM_REQUIRE(dim == 1 || dim == out_shape[d], RM_INVALID_ARG, "fused_elementwise: input %u dim %u (%llu) is not broadcast-compatible with %llu",
                 k, d, (unsigned long long)dim, (unsigned long long)out_shape[d]);
      in_stride[k][d] = dim == 1 ? 0 : s;
      s *= dim;
      same = same && dim == out_shape[d];
    }
    if (elems == 1 && len != 1) scalar_mask |= 1u << k;
    else if (!same) flat = false;
  }

  // outputs
  std::vector<void*> out_ptr(prog.n_outputs);
  for (uint32_t k = 0; k < prog.n_outputs; ++k) RM_TRY(alloc_tensor(p, out_shape, rank, &outs[k], &out_ptr[k]));

  Kernel kern;
  rm_status st;
  std::vector<void*> args;
  for (uint32_t k = 0; k < n_inputs; ++k) args.push_back(&in_ptr[k]);
  for (uint32_t k = 0; k < prog.n_outputs; ++k) args.push_back(&out_ptr[k]);

  if (flat) {
    const std::string k2 = "ewF|" + std::to_string(scalar_mask) + "|" + key;
    st = get_kernel(p, k2, "rm_fused_ew", [&] { return emit_elementwise_cuda(prog, EwVariant::Flat, scalar_mask); }, &kern);
    if (st == RM_OK) {
      unsigned long long n = len;
      args.push_back(&n);
      const uint64_t vec = p->precision == RM_F64 ? 4 : 8;
      const uint64_t nvec = len / vec;
      const uint64_t blocks = std::max<uint64_t>(1, (nvec + 511) / 512);
      st = launch(p, kern, dim3((unsigned)blocks), dim3(256), args.data());
    }
  } else {
    // coalesce dims: drop size-1 dims, merge neighbours whose strides chain for every input
    std::vector<uint64_t> shp;
    std::vector<std::vector<uint64_t>> str(n_inputs);
    for (uint32_t d = 0; d < rank; ++d) {
      if (out_shape[d] == 1) continue;
      bool merge = !shp.empty();
      if (merge)
        for (uint32_t k = 0; k < n_inputs && merge; ++k) merge = in_stride[k][d] == str[k].back() * shp.back();
      if (merge) shp.back() *= out_shape[d];
      else {
        shp.push_back(out_shape[d]);
        for (uint32_t k = 0; k < n_inputs; ++k) str[k].push_back(in_stride[k][d]);
      }
    }
    if (shp.empty()) { shp.push_back(1); for (uint32_t k = 0; k < n_inputs; ++k) str[k].push_back(0); }
    if (shp.size() > 6) {
      for (uint32_t k = 0; k < prog.n_outputs; ++k) rm_free(p, &outs[k]);
      return fail(RM_UNSUPPORTED, "fused_elementwise: broadcast pattern needs %zu dims after coalescing (max 6)", shp.size());
    }
    const std::string k2 = "ewB|" + key;
    st = get_kernel(p, k2, "rm_fused_ew", [&] { return emit_elementwise_cuda(prog, EwVariant::Broadcast, 0); }, &kern);
    if (st == RM_OK) {
      // BParams { u64 len; u32 rank; u32 pad; u64 shape[6]; u64 stride[MAXIN][6]; }
      std::vector<uint64_t> params(2 + 6 + (size_t)n_inputs * 6, 0);
      params[0] = len;
      uint32_t rk[2] = {(uint32_t)shp.size(), 0};
      memcpy(&params[1], rk, 8);
      for (size_t d = 0; d < shp.size(); ++d) params[2 + d] = shp[d];
      for (uint32_t k = 0; k < n_inputs; ++k)
        for (size_t d = 0; d < shp.size(); ++d) params[8 + (size_t)k * 6 + d] = str[k][d];
      args.push_back(params.data());
      const uint64_t blocks = std::min<uint64_t>((len + 255) / 256, (uint64_t)p->prop.multiProcessorCount * 32);
      st = launch(p, kern, dim3((unsigned)std::max<uint64_t>(blocks, 1)), dim3(256), args.data());
    }
  }
  if (st != RM_OK) {
    std::string msg = last_error();
    for (uint32_t k = 0; k < prog.n_outputs; ++k) rm_free(p, &outs[k]);
    set_error("%s", msg.c_str());
  }
  return st;
}

rm_status run_reduction_program(rm_provider* p, const ReductionProgram& prog, const std::string& key, RedOp op,
                                RedLayout layout, const rm_handle* inputs, uint32_t n_inputs,
                                const uint64_t* out_shape, uint32_t rank, uint64_t reduce_len,
                                uint64_t num_slices, uint64_t inner, int use_div, double factor, rm_handle* out, const P2PPublish* publish, double param0) {
  RM_REQUIRE(n_inputs == prog.n_inputs, RM_INVALID_ARG, "fused_reduction: shader declares %u inputs, got %u", prog.n_inputs, n_inputs);
  RM_REQUIRE((p->precision == RM_F64) == (prog.scalar_ty == "f64"), RM_INVALID_ARG,
             "fused_reduction: shader scalar type %s does not match provider precision", prog.scalar_ty.c_str());
  RM_REQUIRE(shape_elems(out_shape, rank) == num_slices, RM_INVALID_ARG, "fused_reduction: output shape does not hold %llu slices",
             (unsigned long long)num_slices);
  RM_REQUIRE(num_slices > 0, RM_INVALID_ARG, "fused_reduction: zero slices");
  std::vector<void*> in_ptr(n_inputs);
  for (uint32_t k = 0; k < n_inputs; ++k) {
    uint64_t elems = 0;
    RM_TRY(resolve(p, &inputs[k], &in_ptr[k], &elems));
    RM_REQUIRE(elems == reduce_len * num_slices, RM_INVALID_ARG, "fused_reduction: input %u has %llu elements, expected %llu x %llu", k,
               (unsigned long long)elems, (unsigned long long)reduce_len, (unsigned long long)num_slices);
  }
  if (num_slices == 1) layout = RedLayout::Contig;
  const uint64_t vec = p->precision == RM_F64 ? 4 : 8;
  if (layout == RedLayout::Strided && inner >= 2 && (inner & (inner - 1)) == 0 && inner <= 32 * vec && num_slices % inner == 0 &&
      (inner * reduce_len) % vec == 0 && inner * reduce_len >= 256 * vec * 8 && num_slices / inner <= 65535 && !getenv("RUNMAT_B200_RED_NO_INTERLEAVED"))
    layout = RedLayout::Interleaved;
  RM_REQUIRE(!publish || num_slices == 1, RM_INVALID_ARG, "fused_reduction_allreduce: only scalar ('all') reductions are exchanged");
  void* out_ptr = nullptr;
  RM_TRY(alloc_tensor(p, out_shape, rank, out, &out_ptr));

  const uint64_t sms = (uint64_t)p->prop.multiProcessorCount;
  Kernel kern;
  const std::string k2 = std::string(layout == RedLayout::Contig ? "redC|" : layout == RedLayout::Strided ? "redS|" : "redI|") + std::to_string((int)op) + "|" + key;
  rm_status st = get_kernel(p, k2, "rm_fused_red", [&] { return emit_reduction_cuda(prog, op, layout); }, &kern);
  if (st == RM_OK) {
    dim3 grid, block;
    uint64_t partial_elems = 0;
    int vec_ok = 0;
    unsigned bps = 1, sl = 1;
    if (layout == RedLayout::Contig) {
      vec_ok = (reduce_len % vec == 0 || num_slices == 1) ? 1 : 0;
      uint64_t threads = 256;
      while (threads > 32 && threads * vec / 2 >= reduce_len) threads >>= 1;
      if (num_slices < 2 * sms && num_slices <= FusedCache::kTicketCap) {
        int bpsm = 4;  // r01 sweep: 4 CTAs/SM beats 8 and 16 (the kernel is ~50 us: per-block tail and finish costs dominate)
        if (const char* e = getenv("RUNMAT_B200_RED_BPSM")) { int v = atoi(e); if (v >= 1 && v <= 64) bpsm = v; }
        const uint64_t want = (sms * (uint64_t)bpsm + num_slices - 1) / num_slices;
        const uint64_t max_bps = std::max<uint64_t>(1, reduce_len / (threads * vec * 4));
        bps = (unsigned)std::min<uint64_t>(std::min(want, max_bps), 4096);
      }
      RM_REQUIRE(num_slices * bps < (1ull << 31), RM_UNSUPPORTED, "fused_reduction: too many slices");
      grid = dim3((unsigned)(num_slices * bps));
      block = dim3((unsigned)threads);
      partial_elems = bps > 1 ? (uint64_t)bps * num_slices : 0;
    } else if (layout == RedLayout::Interleaved) {
      // one flat vector stream per [inner x len] block; one resident wave (4 CTAs/SM) shared by the blocks
      const uint64_t blocks = num_slices / inner, nvec = inner * reduce_len / vec;
      uint64_t gx = std::max<uint64_t>(1, (sms * 4 + blocks - 1) / blocks);
      gx = std::min<uint64_t>(gx, std::max<uint64_t>(1, nvec / (256 * 2)));
      grid = dim3((unsigned)gx, (unsigned)blocks);
      block = dim3(256);
      partial_elems = gx > 1 ? gx * num_slices : 0;
    } else {
      // sl adjacent slices per CTA (power of two), 256/sl row lanes; split rows over gridDim.y until the SMs are full
      while (sl < 256 && sl < num_slices) sl <<= 1;
      const uint64_t rl = 256 / sl;
      const uint64_t gx = (num_slices + sl - 1) / sl;
      uint64_t chunks = 1;
      if (gx < sms * 8 && gx <= FusedCache::kTicketCap)
        chunks = std::min<uint64_t>(std::min<uint64_t>((sms * 8 + gx - 1) / gx, 65535), std::max<uint64_t>(1, reduce_len / (rl * 8)));
      RM_REQUIRE(gx < (1ull << 31), RM_UNSUPPORTED, "fused_reduction: too many slices");
      grid = dim3((unsigned)gx, (unsigned)chunks);
      block = dim3(256);
      partial_elems = chunks > 1 ? chunks * num_slices : 0;
    }
    // held across ensure -> pointer capture -> launch: a concurrent host thread may not re-allocate the scratch in between
    std::lock_guard<std::mutex> scratch_lock(p->scratch_mu);
    st = ensure_scratch(p, partial_elems * (sizeof(double) + sizeof(uint32_t)) + 256);
    if (st == RM_OK) {
      double* partial = (double*)p->reduce_scratch;
      uint32_t* pflags = (uint32_t*)((char*)p->reduce_scratch + ((partial_elems * sizeof(double) + 255) / 256) * 256);
      uint32_t* tickets = (uint32_t*)p->fused->tickets;
      unsigned long long len_arg = reduce_len, slices_arg = num_slices, inner_arg = inner ? inner : 1;
      std::vector<void*> args;
      for (uint32_t k = 0; k < n_inputs; ++k) args.push_back(&in_ptr[k]);
      args.push_back(&out_ptr);
      args.push_back(&partial);
      args.push_back(&pflags);
      args.push_back(&tickets);
      args.push_back(&len_arg);
      args.push_back(&slices_arg);
      args.push_back(&vec_ok);
      args.push_back(&use_div);
      args.push_back(&factor);
      args.push_back(&bps);
      args.push_back(&inner_arg);
      args.push_back(&sl);
      // fused peer-memory publish of the scalar result (comm.cu); n == 0 disables the tail
      void* const* pub_peers = publish ? publish->peers : nullptr;
      uint32_t pub_n = publish ? publish->n : 0, pub_rank = publish ? publish->rank : 0;
      unsigned long long pub_step = publish ? publish->step : 0;
      args.push_back(&pub_peers);
      args.push_back(&pub_n);
      args.push_back(&pub_rank);
      args.push_back(&pub_step);
      void* pub_prev_dst = publish ? publish->prev_dst : nullptr;
      unsigned long long pub_prev_step1 = publish ? publish->prev_step1 : 0;
      int* pub_err = publish ? publish->err : nullptr;
      args.push_back(&pub_prev_dst);
      args.push_back(&pub_prev_step1);
      args.push_back(&pub_err);
      args.push_back(&param0);  // free scalar of the value expression (`p0`), e.g. the strike of the payoff reduction
      st = launch(p, kern, grid, block, args.data());
    }
  }
  if (st != RM_OK) {
    std::string msg = last_error();
    rm_free(p, out);
    set_error("%s", msg.c_str());
  }
  return st;
}

}  // namespace rm
