// jit.cu - generic per-quadrature-point operators compiled at run time (NVRTC) for sm_100a.
//
// replaces: the open-ended `external_function(derivatives)(*operands)` protocol of
// src/dolfinx_external_operator/external_operator.py:432 for models that are NOT one of the hard-wired
// kernels: the user hands in the CUDA C++ text of one function template (see include/eo_dual.h) and gets the
// value, first and second derivatives with respect to any operand by forward-mode dual numbers - what the
// reference obtains from JAX / torch.func (README.md:16-25, demo_mc:555, demo_hyperelasticity.py:429-456).
//
// NVRTC (libnvrtc.so.12) is loaded with dlopen on first use so that the library itself keeps no link-time
// dependency on it; the CUBIN is loaded with cudaLibraryLoadData and launched with cudaLaunchKernel on the
// context's compute stream, host-side arguments go through the same chunked pipeline as every other entry point.
#include <dlfcn.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <string>

#include "eo_common.cuh"
#include "eo_jit_device.cuh"  // eo_jit_args (the same text NVRTC sees)

static const char* k_hdr_dual =
#include "eo_dual.inc"
    ;
static const char* k_hdr_device =
#include "eo_jit_device.inc"
    ;
static const char* k_hdr_tab =
#include "tab_core.inc"
    ;
#include "tab_core.cuh"  // tab_tables (sizes of the element the fused variants are specialised for)

// ---------------------------------------------------------------------------------------------
// NVRTC through dlopen
// ---------------------------------------------------------------------------------------------
namespace {
typedef struct _nvrtcProgram* nvrtcProgram;
struct nvrtc_api {
  void* h = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  int (*CompileProgram)(nvrtcProgram, int, const char* const*);
  int (*DestroyProgram)(nvrtcProgram*);
  int (*GetCUBINSize)(nvrtcProgram, size_t*);
  int (*GetCUBIN)(nvrtcProgram, char*);
  int (*GetProgramLogSize)(nvrtcProgram, size_t*);
  int (*GetProgramLog)(nvrtcProgram, char*);
  const char* (*GetErrorString)(int);
  int (*Version)(int*, int*);
  std::string err;
  int major = 0, minor = 0;
};

nvrtc_api* nvrtc() {
  static nvrtc_api api;
  static std::once_flag once;
  std::call_once(once, [] {
    // The toolkit's NVRTC first, by path: a bare soname would resolve to whichever libnvrtc.so.12 the process has
    // already loaded (PyTorch ships 12.8, whose ptxas rejects the 256-bit vector accesses of sm_100).
    const char* names[] = {getenv("EO_NVRTC_LIB"),
                           "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so",
                           "libnvrtc.so.12",
                           "libnvrtc.so.13",
                           "libnvrtc.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      api.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (api.h) break;
    }
    if (!api.h) {
      api.err = "libnvrtc.so.12 not found (set EO_NVRTC_LIB to its path)";
      return;
    }
#define EO_SYM(field, name)                                   \
  *(void**)(&api.field) = dlsym(api.h, name);                 \
  if (!api.field) {                                           \
    api.err = std::string("libnvrtc lacks ") + name;          \
    api.h = nullptr;                                          \
    return;                                                   \
  }
    EO_SYM(CreateProgram, "nvrtcCreateProgram");
    EO_SYM(CompileProgram, "nvrtcCompileProgram");
    EO_SYM(DestroyProgram, "nvrtcDestroyProgram");
    EO_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    EO_SYM(GetCUBIN, "nvrtcGetCUBIN");
    EO_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    EO_SYM(GetProgramLog, "nvrtcGetProgramLog");
    EO_SYM(GetErrorString, "nvrtcGetErrorString");
    EO_SYM(Version, "nvrtcVersion");
#undef EO_SYM
    api.Version(&api.major, &api.minor);
  });
  return &api;
}

struct jit_variant {
  std::string cubin;
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kernel = nullptr;
  int order = 0, da = 0, db = 0;
  int out_width = 0;  // doubles per point of the primary output
  int tile = 0;       // staged variant: points per CTA (0: direct variant)
  int ppt = 0;        // several-points-per-thread variant: points per thread
  size_t smem = 0;    // staged variant: dynamic shared memory per CTA
};
}  // namespace

struct eo_jit {
  eo_ctx* ctx = nullptr;  // may be nullptr: compile-only object (no GPU needed)
  std::string source, entry;
  int n_operands = 0, n_state = 0, n_aux = 0, n_params = 0, out_size = 0, fmad = 1;
  int operand_size[EO_JIT_MAX_ARGS] = {}, state_size[EO_JIT_MAX_ARGS] = {}, aux_size[EO_JIT_MAX_ARGS] = {};
  std::map<std::string, jit_variant> variants;
  std::string log;  // last compile log
  char err[512] = {0};
};

static int jit_fail(eo_jit* m, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(m->err, sizeof m->err, fmt, ap);
  va_end(ap);
  if (m->ctx) snprintf(m->ctx->err, sizeof m->ctx->err, "%s", m->err);
  return code;
}

static std::string int_list(const int* v, int n) {
  std::string s = "{";
  for (int i = 0; i < n; ++i) s += (i ? "," : "") + std::to_string(v[i]);
  if (n == 0) s += "0";
  return s + "}";
}

// tabulation source of one operand in a fused variant
struct jit_tab_sig {
  int gdim, bs, nb, kind;
};

// The translation unit handed to NVRTC for one derivative multi-index (fused: nq > 0 and one jit_tab_sig per operand).
// tile > 0: the staged (TMA bulk copy) variant with `tile` points per CTA.
// ppt > 1: the several-points-per-thread variant (scalar-sized models).
static std::string jit_program(const eo_jit* m, int order, int da, int db, int nq = 0, const jit_tab_sig* sig = nullptr, int tile = 0,
                               int ppt = 0) {
  int nin = 0, nst = 0, naux = 0;
  for (int i = 0; i < m->n_operands; ++i) nin += m->operand_size[i];
  for (int i = 0; i < m->n_state; ++i) nst += m->state_size[i];
  for (int i = 0; i < m->n_aux; ++i) naux += m->aux_size[i];
  std::string s;
  if (sig) s += "#define EO_JIT_FUSED 1\n#define EO_JIT_N_TABLES " + std::to_string(m->n_operands) + "\n";
  if (tile) s += "#define EO_JIT_STAGED 1\n";
  if (ppt > 1) s += "#define EO_JIT_PPT 1\n";
  s += "#include \"eo_jit_device.cuh\"\n";
  s += "#line 1 \"model.cu\"\n";
  s += m->source;
  s += "\n#line 1 \"eo_jit_entry.cu\"\n";
  s += "struct eo_jit_spec {\n";
  s += "  static constexpr int N_OPERANDS = " + std::to_string(m->n_operands) + ", N_STATE = " + std::to_string(m->n_state) +
       ", N_AUX = " + std::to_string(m->n_aux) + ";\n";
  s += "  static constexpr int NIN = " + std::to_string(nin) + ", NST = " + std::to_string(nst) +
       ", NOUT = " + std::to_string(m->out_size) + ", NAUX = " + std::to_string(naux) + ";\n";
  if (tile) s += "  static constexpr int TILE = " + std::to_string(tile) + ";\n";
  if (ppt > 1) s += "  static constexpr int PPT = " + std::to_string(ppt) + ";\n";
  s += "  static constexpr int ORDER = " + std::to_string(order) + ", DA = " + std::to_string(da) + ", DB = " + std::to_string(db) + ";\n";
  s += "  static constexpr int op_size(int k) { constexpr int t[] = " + int_list(m->operand_size, m->n_operands) + "; return t[k]; }\n";
  s += "  static constexpr int st_size(int k) { constexpr int t[] = " + int_list(m->state_size, m->n_state) + "; return t[k]; }\n";
  s += "  static constexpr int aux_size(int k) { constexpr int t[] = " + int_list(m->aux_size, m->n_aux) + "; return t[k]; }\n";
  if (sig) {
    int g[EO_JIT_MAX_ARGS], b[EO_JIT_MAX_ARGS], nb[EO_JIT_MAX_ARGS], kd[EO_JIT_MAX_ARGS];
    for (int i = 0; i < m->n_operands; ++i) g[i] = sig[i].gdim, b[i] = sig[i].bs, nb[i] = sig[i].nb, kd[i] = sig[i].kind;
    s += "  static constexpr int NQ = " + std::to_string(nq) + ";\n";
    s += "  static constexpr int tab_gdim(int k) { constexpr int t[] = " + int_list(g, m->n_operands) + "; return t[k]; }\n";
    s += "  static constexpr int tab_bs(int k) { constexpr int t[] = " + int_list(b, m->n_operands) + "; return t[k]; }\n";
    s += "  static constexpr int tab_nb(int k) { constexpr int t[] = " + int_list(nb, m->n_operands) + "; return t[k]; }\n";
    s += "  static constexpr int tab_kind(int k) { constexpr int t[] = " + int_list(kd, m->n_operands) + "; return t[k]; }\n";
  }
  s += "  template <class T> static __device__ __forceinline__ void call(const T* x, const double* s, const double* prm, T* y, T* w) {\n";
  s += "    " + m->entry + "<T>(x, s, prm, y, w);\n  }\n};\n";
  // occupancy hint: the fused kernels hide gather latency with resident warps (tab_vm_kernel uses 4 CTAs/SM too)
  int min_blocks = sig ? 4 : 1;
  if (const char* e = getenv("EO_JIT_MIN_BLOCKS")) min_blocks = atoi(e) > 0 ? atoi(e) : min_blocks;
  if (tile) {
    s += "extern \"C\" __global__ void __launch_bounds__(" + std::to_string(tile) + ") eo_jit_entry(const __grid_constant__ eo_jit_args a) {\n";
    s += "  eo_jitd::run_staged<eo_jit_spec>(a);\n}\n";
    return s;
  }
  if (ppt > 1) {
    s += "extern \"C\" __global__ void __launch_bounds__(256) eo_jit_entry(const __grid_constant__ eo_jit_args a) {\n";
    s += "  eo_jitd::run_ppt<eo_jit_spec>(a);\n}\n";
    return s;
  }
  s += "extern \"C\" __global__ void __launch_bounds__(256, " + std::to_string(min_blocks) + ") eo_jit_entry(const __grid_constant__ " +
       std::string(sig ? "eo_jit_fused_args" : "eo_jit_args") + " a) {\n";
  s += std::string("  eo_jitd::run<eo_jit_spec, ") + (sig ? "eo_jit_fused_args" : "eo_jit_args") + ">(a);\n}\n";
  return s;
}

static int jit_multi_index(eo_jit* m, const int* derivatives, int& order, int& da, int& db) {
  order = 0, da = 0, db = 0;
  for (int i = 0; i < m->n_operands; ++i) {
    const int d = derivatives ? derivatives[i] : 0;
    if (d < 0) return jit_fail(m, EO_ERR_INVALID, "eo_jit: negative derivative order");
    for (int r = 0; r < d; ++r) {
      if (order == 0) da = i;
      if (order == 1) db = i;
      ++order;
    }
  }
  if (order > 2) return jit_fail(m, EO_ERR_UNSUPPORTED, "eo_jit: derivative order %d > 2 is not implemented", order);
  return EO_OK;
}

static int jit_compile(eo_jit* m, int order, int da, int db, jit_variant** out, int nq = 0, const jit_tab_sig* sig = nullptr,
                       int tile = 0, int ppt = 0) {
  std::string key = std::to_string(order) + ":" + std::to_string(da) + ":" + std::to_string(db);
  if (tile) key += ":T" + std::to_string(tile);
  if (ppt > 1) key += ":P" + std::to_string(ppt);
  if (sig) {
    key += ":F" + std::to_string(nq);
    for (int i = 0; i < m->n_operands; ++i)
      key += "/" + std::to_string(sig[i].gdim) + "," + std::to_string(sig[i].bs) + "," + std::to_string(sig[i].nb) + "," + std::to_string(sig[i].kind);
  }
  auto it = m->variants.find(key);
  if (it != m->variants.end()) {
    *out = &it->second;
    return EO_OK;
  }
  nvrtc_api* rt = nvrtc();
  if (!rt->h) return jit_fail(m, EO_ERR_UNSUPPORTED, "eo_jit: %s", rt->err.c_str());
  const std::string prog_text = jit_program(m, order, da, db, nq, sig, tile, ppt);
  // on-disk CUBIN cache (EO_JIT_CACHE_DIR): keyed by the full translation unit, both headers, the options and the
  // NVRTC version, so a process restart does not pay the 0.1-3 s compilation again
  std::string cache_path;
  if (const char* dir = getenv("EO_JIT_CACHE_DIR")) {
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](const char* p, size_t n) {
      for (size_t i = 0; i < n; ++i) h = (h ^ (unsigned char)p[i]) * 1099511628211ull;
    };
    mix(prog_text.data(), prog_text.size());
    mix(k_hdr_device, strlen(k_hdr_device));
    mix(k_hdr_dual, strlen(k_hdr_dual));
    mix(k_hdr_tab, strlen(k_hdr_tab));
    const int sig_opts[3] = {m->fmad, rt->major, rt->minor};
    mix(reinterpret_cast<const char*>(sig_opts), sizeof sig_opts);
    char name[64];
    snprintf(name, sizeof name, "/eo_jit_%016llx.cubin", h);
    cache_path = std::string(dir) + name;
  }
  nvrtcProgram prog = nullptr;
  const char* hdr_src[3] = {k_hdr_device, k_hdr_dual, k_hdr_tab};
  const char* hdr_name[3] = {"eo_jit_device.cuh", "eo_dual.h", "tab_core.cuh"};
  jit_variant v;
  bool cached = false;
  if (!cache_path.empty()) {
    if (FILE* f = fopen(cache_path.c_str(), "rb")) {
      fseek(f, 0, SEEK_END);
      const long sz = ftell(f);
      fseek(f, 0, SEEK_SET);
      if (sz > 64) {
        v.cubin.assign(size_t(sz), '\0');
        cached = fread(&v.cubin[0], 1, size_t(sz), f) == size_t(sz) && v.cubin.compare(0, 4, "\x7f" "ELF") == 0;
      }
      fclose(f);
      if (cached) m->log = "loaded from " + cache_path;
    }
  }
  int rc = 0;
  if (!cached) {
  rc = rt->CreateProgram(&prog, prog_text.c_str(), "eo_jit_entry.cu", 3, hdr_src, hdr_name);
  if (rc) return jit_fail(m, EO_ERR_CUDA, "nvrtcCreateProgram: %s", rt->GetErrorString(rc));
  // 256-bit ld/st.global.v4.f64 need the CUDA >= 12.9 ptxas; older NVRTCs get 128-bit accesses
  const bool v4 = rt->major > 12 || (rt->major == 12 && rt->minor >= 9);
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device",
                        m->fmad ? "--fmad=true" : "--fmad=false", v4 ? "-DEO_JIT_MAX_VEC=4" : "-DEO_JIT_MAX_VEC=2"};
  rc = rt->CompileProgram(prog, 6, opts);
  size_t ls = 0;
  rt->GetProgramLogSize(prog, &ls);
  m->log.assign(ls ? ls : 1, '\0');
  if (ls) rt->GetProgramLog(prog, &m->log[0]);
  while (!m->log.empty() && m->log.back() == '\0') m->log.pop_back();
  if (rc) {
    rt->DestroyProgram(&prog);
    return jit_fail(m, EO_ERR_INVALID, "eo_jit: compilation of '%s' (derivatives %s) failed: %s; see eo_jit_log", m->entry.c_str(),
                    key.c_str(), rt->GetErrorString(rc));
  }
  size_t cs = 0;
  rt->GetCUBINSize(prog, &cs);
  v.cubin.assign(cs, '\0');
  rt->GetCUBIN(prog, &v.cubin[0]);
  rt->DestroyProgram(&prog);
  if (!cache_path.empty()) {  // write-then-rename: concurrent ranks compile the same model at start-up
    const std::string tmp = cache_path + "." + std::to_string((long)getpid());
    if (FILE* f = fopen(tmp.c_str(), "wb")) {
      const bool ok = fwrite(v.cubin.data(), 1, v.cubin.size(), f) == v.cubin.size();
      fclose(f);
      if (!ok || rename(tmp.c_str(), cache_path.c_str()) != 0) remove(tmp.c_str());
    }
  }
  }  // !cached
  if (const char* dump = getenv("EO_JIT_DUMP_DIR")) {  // for cuobjdump -sass / -res-usage
    std::string name = key;
    for (char& ch : name)
      if (ch == ':' || ch == '/' || ch == ',') ch = '_';
    const std::string path = std::string(dump) + "/" + m->entry + "_" + name + ".cubin";
    if (FILE* f = fopen(path.c_str(), "wb")) {
      fwrite(v.cubin.data(), 1, v.cubin.size(), f);
      fclose(f);
    }
  }
  v.order = order, v.da = da, v.db = db;
  v.out_width = m->out_size;
  if (order >= 1) v.out_width *= m->operand_size[da];
  if (order >= 2) v.out_width *= m->operand_size[db];
  v.tile = tile;
  v.ppt = ppt;
  if (tile) {
    int doubles = v.out_width + m->out_size;
    for (int i = 0; i < m->n_operands; ++i) doubles += m->operand_size[i];
    for (int i = 0; i < m->n_state; ++i) doubles += m->state_size[i];
    for (int i = 0; i < m->n_aux; ++i) doubles += m->aux_size[i];
    v.smem = 128 + size_t(tile) * 8 * doubles;
  }
  if (m->ctx) {
    cudaError_t e = cudaLibraryLoadData(&v.lib, v.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) return jit_fail(m, EO_ERR_CUDA, "cudaLibraryLoadData: %s", cudaGetErrorString(e));
    e = cudaLibraryGetKernel(&v.kernel, v.lib, "eo_jit_entry");
    if (e != cudaSuccess) return jit_fail(m, EO_ERR_CUDA, "cudaLibraryGetKernel: %s", cudaGetErrorString(e));
    if (v.smem > 48 * 1024) {
      e = cudaFuncSetAttribute((const void*)v.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem);
      if (e != cudaSuccess) return jit_fail(m, EO_ERR_CUDA, "cudaFuncSetAttribute(staged kernel, %zu B): %s", v.smem, cudaGetErrorString(e));
    }
  }
  auto ins = m->variants.emplace(key, std::move(v));
  *out = &ins.first->second;
  return EO_OK;
}

// Points per CTA (= threads) of the staged variant for one derivative multi-index: the largest power of two <= 1024
// whose tile (every per-point array of the call) fits in 32 KB - several CTAs per SM, so that the load, compute and
// store phases of different CTAs overlap - but no smaller than 32; 0 when even 32 points do not fit in 96 KB.
static int jit_tile_points(const eo_jit* m, int order, int da, int db) {
  int doubles = m->out_size * (order >= 1 ? m->operand_size[da] : 1) * (order >= 2 ? m->operand_size[db] : 1) + m->out_size;
  for (int i = 0; i < m->n_operands; ++i) doubles += m->operand_size[i];
  for (int i = 0; i < m->n_state; ++i) doubles += m->state_size[i];
  for (int i = 0; i < m->n_aux; ++i) doubles += m->aux_size[i];
  for (int t = 1024; t >= 32; t /= 2)
    if (128 + size_t(t) * 8 * doubles <= (t == 32 ? 96u : 32u) * 1024) return t;
  return 0;
}

// Points per thread for scalar-sized models (<= 4 doubles read, <= 8 written per point): >= 32 bytes of loads per thread.
static int jit_points_per_thread(const eo_jit* m, int out_width) {
  int in = 0;
  for (int i = 0; i < m->n_operands; ++i) in += m->operand_size[i];
  for (int i = 0; i < m->n_state; ++i) in += m->state_size[i];
  int aux = 0;
  for (int i = 0; i < m->n_aux; ++i) aux += m->aux_size[i];
  if (in > 4 || out_width + m->out_size + aux > 8) return 1;
  return in == 1 ? 4 : 2;
}

// Is the staged variant worth it?  Per-thread accesses are strided (0.15-0.68 of the HBM roofline measured) when a
// point has an odd number (>= 3) of components; scalars are contiguous across a warp and even counts vectorise.
static bool jit_has_odd_array(const eo_jit* m, int out_width, int order) {
  auto bad = [](int s) { return s >= 3 && (s % 2) != 0; };
  bool odd = bad(out_width) || (order >= 1 && bad(m->out_size));
  for (int i = 0; i < m->n_operands; ++i) odd = odd || bad(m->operand_size[i]);
  for (int i = 0; i < m->n_state; ++i) odd = odd || bad(m->state_size[i]);
  for (int i = 0; i < m->n_aux; ++i) odd = odd || bad(m->aux_size[i]);
  return odd;
}

extern "C" {

int eo_jit_create(eo_ctx* ctx, const eo_jit_desc* d, eo_jit** out) {
  if (!out) return eo_fail(ctx, EO_ERR_INVALID, "eo_jit_create: out is NULL");
  *out = nullptr;
  if (!d || !d->source || !d->entry) return eo_fail(ctx, EO_ERR_INVALID, "eo_jit_create: NULL descriptor / source / entry");
  if (d->n_operands < 1 || d->n_operands > EO_JIT_MAX_ARGS || d->n_state < 0 || d->n_state > EO_JIT_MAX_ARGS ||
      d->n_aux < 0 || d->n_aux > EO_JIT_MAX_ARGS || d->n_params < 0 || d->n_params > EO_JIT_MAX_PARAMS)
    return eo_fail(ctx, EO_ERR_INVALID, "eo_jit_create: operand / state / aux / parameter count out of range");
  if (d->out_size < 1 || d->out_size > 64) return eo_fail(ctx, EO_ERR_INVALID, "eo_jit_create: out_size must be in [1, 64]");
  eo_jit* m = new eo_jit();
  m->ctx = ctx;
  m->source = d->source;
  m->entry = d->entry;
  m->n_operands = d->n_operands, m->n_state = d->n_state, m->n_aux = d->n_aux, m->n_params = d->n_params;
  m->out_size = d->out_size;
  m->fmad = d->fmad ? 1 : 0;
  bool ok = true;
  for (int i = 0; i < d->n_operands; ++i) ok = ok && (m->operand_size[i] = d->operand_size[i]) >= 1 && d->operand_size[i] <= 16;
  for (int i = 0; i < d->n_state; ++i) ok = ok && (m->state_size[i] = d->state_size[i]) >= 1 && d->state_size[i] <= 64;
  for (int i = 0; i < d->n_aux; ++i) ok = ok && (m->aux_size[i] = d->aux_size[i]) >= 1 && d->aux_size[i] <= 64;
  if (!ok) {
    delete m;
    return eo_fail(ctx, EO_ERR_INVALID, "eo_jit_create: a component count is out of range (operands 1..16, state/aux 1..64)");
  }
  *out = m;
  return EO_OK;
}

int eo_jit_destroy(eo_jit* m) {
  if (!m) return EO_OK;
  if (m->ctx) cudaStreamSynchronize(m->ctx->s_cmp);
  for (auto& kv : m->variants)
    if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
  delete m;
  return EO_OK;
}

int eo_jit_nvrtc_version(void) {
  nvrtc_api* rt = nvrtc();
  return rt->h ? rt->major * 1000 + rt->minor * 10 : EO_ERR_UNSUPPORTED;
}

const char* eo_jit_log(const eo_jit* m) { return m ? m->log.c_str() : ""; }
const char* eo_jit_last_error(const eo_jit* m) { return m ? m->err : "eo_jit: NULL handle"; }

int eo_jit_compile(eo_jit* m, const int* derivatives, size_t* cubin_bytes) {
  if (!m) return EO_ERR_INVALID;
  int order, da, db;
  int rc = jit_multi_index(m, derivatives, order, da, db);
  if (rc) return rc;
  jit_variant* v = nullptr;
  rc = jit_compile(m, order, da, db, &v);
  if (rc) return rc;
  if (cubin_bytes) *cubin_bytes = v->cubin.size();
  return EO_OK;
}

int eo_jit_compile_staged(eo_jit* m, const int* derivatives, int* tile_points, size_t* cubin_bytes) {
  if (!m) return EO_ERR_INVALID;
  int order, da, db;
  int rc = jit_multi_index(m, derivatives, order, da, db);
  if (rc) return rc;
  const int tile = jit_tile_points(m, order, da, db);
  if (tile_points) *tile_points = tile;
  if (!tile) return jit_fail(m, EO_ERR_UNSUPPORTED, "eo_jit_compile_staged: a 32-point tile of this model does not fit in shared memory");
  jit_variant* v = nullptr;
  rc = jit_compile(m, order, da, db, &v, 0, nullptr, tile);
  if (rc) return rc;
  if (cubin_bytes) *cubin_bytes = v->cubin.size();
  return EO_OK;
}

int eo_jit_compile_ppt(eo_jit* m, const int* derivatives, int* points_per_thread, size_t* cubin_bytes) {
  if (!m) return EO_ERR_INVALID;
  int order, da, db;
  int rc = jit_multi_index(m, derivatives, order, da, db);
  if (rc) return rc;
  const int w = m->out_size * (order >= 1 ? m->operand_size[da] : 1) * (order >= 2 ? m->operand_size[db] : 1);
  const int ppt = jit_points_per_thread(m, w);
  if (points_per_thread) *points_per_thread = ppt;
  if (ppt <= 1) return jit_fail(m, EO_ERR_UNSUPPORTED, "eo_jit_compile_ppt: this model is not scalar-sized (one point per thread is used)");
  jit_variant* v = nullptr;
  rc = jit_compile(m, order, da, db, &v, 0, nullptr, 0, ppt);
  if (rc) return rc;
  if (cubin_bytes) *cubin_bytes = v->cubin.size();
  return EO_OK;
}

int eo_jit_cubin(eo_jit* m, const int* derivatives, void* buf, size_t buf_bytes) {
  if (!m) return EO_ERR_INVALID;
  int order, da, db;
  int rc = jit_multi_index(m, derivatives, order, da, db);
  if (rc) return rc;
  jit_variant* v = nullptr;
  rc = jit_compile(m, order, da, db, &v);
  if (rc) return rc;
  if (!buf || buf_bytes < v->cubin.size()) return jit_fail(m, EO_ERR_INVALID, "eo_jit_cubin: buffer too small (%zu needed)", v->cubin.size());
  memcpy(buf, v->cubin.data(), v->cubin.size());
  return EO_OK;
}

int eo_jit_out_width(eo_jit* m, const int* derivatives) {
  if (!m) return EO_ERR_INVALID;
  int order, da, db;
  int rc = jit_multi_index(m, derivatives, order, da, db);
  if (rc) return rc;
  int w = m->out_size;
  if (order >= 1) w *= m->operand_size[da];
  if (order >= 2) w *= m->operand_size[db];
  return w;
}

int eo_jit_eval(eo_jit* m, const int* derivatives, const double* params, const double* const* operands,
                const double* const* state, double* out, double* value, double* const* aux, int64_t n) {
  if (!m) return EO_ERR_INVALID;
  eo_ctx* ctx = m->ctx;
  if (!ctx) return jit_fail(m, EO_ERR_NO_DEVICE, "eo_jit_eval: this model was created without a context (compile-only)");
  if (n < 0) return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval: n < 0");
  if (!out || !operands || (m->n_state && !state) || (m->n_params && !params))
    return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval: NULL out / operands / state / params");
  int order, da, db;
  int rc = jit_multi_index(m, derivatives, order, da, db);
  if (rc) return rc;
  jit_variant* v = nullptr;
  rc = jit_compile(m, order, da, db, &v);
  if (rc) return rc;
  EO_CUDA(ctx, cudaSetDevice(ctx->device));

  // argument table of the streamed pipeline: operands, state, out, value, aux
  eo_arg args[3 * EO_JIT_MAX_ARGS + 2];
  int na = 0;
  for (int i = 0; i < m->n_operands; ++i) {
    if (!operands[i]) return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval: operand %d is NULL", i);
    args[na++] = {operands[i], size_t(m->operand_size[i]) * 8, false};
  }
  for (int i = 0; i < m->n_state; ++i) {
    if (!state[i]) return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval: state field %d is NULL", i);
    args[na++] = {state[i], size_t(m->state_size[i]) * 8, false};
  }
  const int i_out = na;
  args[na++] = {out, size_t(v->out_width) * 8, true};
  const int i_val = na;
  args[na++] = {order >= 1 ? value : nullptr, size_t(m->out_size) * 8, true};
  const int i_aux = na;
  for (int i = 0; i < m->n_aux; ++i) args[na++] = {aux ? aux[i] : nullptr, size_t(m->aux_size[i]) * 8, true};
  for (int i = 0; i < na; ++i)
    if (args[i].ptr && (reinterpret_cast<uintptr_t>(args[i].ptr) & 7))
      return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval: argument %d is not 8-byte aligned", i);

  // staged (TMA bulk copy) variant: EO_JIT_STAGED=0 never, =1 whenever it fits, default: when a point of some array
  // has an odd number of components (per-thread accesses would be strided)
  jit_variant* vs = nullptr;
  {
    const char* pol = getenv("EO_JIT_STAGED");
    const bool want = pol ? (*pol == '1') : jit_has_odd_array(m, v->out_width, order);
    const int tile = want ? jit_tile_points(m, order, da, db) : 0;
    if (tile && n >= tile) {
      rc = jit_compile(m, order, da, db, &vs, 0, nullptr, tile);
      if (rc) return rc;
    }
    const char* pp = getenv("EO_JIT_PPT");
    const int ppt = (vs || (pp && *pp == '0')) ? 1 : jit_points_per_thread(m, v->out_width);
    if (ppt > 1 && n >= 1024) {  // bulk variant for scalar-sized models: several points per thread
      rc = jit_compile(m, order, da, db, &vs, 0, nullptr, 0, ppt);
      if (rc) return rc;
    }
  }

  eo_jit_args ka;
  memset(&ka, 0, sizeof ka);
  for (int i = 0; i < m->n_params; ++i) ka.prm[i] = params[i];
  auto launch = [&](void** p, int64_t nn, int64_t) -> int {
    bool a16 = true;
    for (int i = 0; i < na; ++i) {
      if (p[i] && (reinterpret_cast<uintptr_t>(p[i]) & (args[i].bpq % 32 == 0 ? 31 : args[i].bpq % 16 == 0 ? 15 : 7)))
        return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval: device argument %d is misaligned for its %zu-byte points", i, args[i].bpq);
      a16 = a16 && !(reinterpret_cast<uintptr_t>(p[i]) & 15);
    }
    // full tiles through the staged kernel (bulk copies need 16-byte aligned addresses) or full groups through the
    // points-per-thread kernel (32-byte aligned wide accesses), the rest directly
    bool a32 = true;
    for (int i = 0; i < na; ++i) a32 = a32 && !(reinterpret_cast<uintptr_t>(p[i]) & 31);
    const int64_t unit = vs ? (vs->tile ? vs->tile : vs->ppt) : 1;
    const int64_t n_staged = (vs && (vs->tile ? a16 : a32)) ? nn / unit * unit : 0;
    for (int pass = 0; pass < 2; ++pass) {
      const int64_t first = pass == 0 ? 0 : n_staged, cnt = pass == 0 ? n_staged : nn - n_staged;
      if (cnt <= 0) continue;
      auto at = [&](int i) -> char* { return p[i] ? (char*)p[i] + size_t(first) * args[i].bpq : nullptr; };
      for (int i = 0; i < m->n_operands; ++i) ka.operand[i] = (const double*)at(i);
      for (int i = 0; i < m->n_state; ++i) ka.state[i] = (const double*)at(m->n_operands + i);
      ka.out = (double*)at(i_out);
      ka.value = (double*)at(i_val);
      for (int i = 0; i < m->n_aux; ++i) ka.aux[i] = (double*)at(i_aux + i);
      ka.n = cnt;
      void* kargs[1] = {&ka};
      cudaError_t e;
      if (pass == 0 && vs->tile)
        e = cudaLaunchKernel((const void*)vs->kernel, dim3(unsigned(cnt / vs->tile)), dim3(vs->tile), kargs, vs->smem, ctx->s_cmp);
      else if (pass == 0)
        e = cudaLaunchKernel((const void*)vs->kernel, dim3(unsigned((cnt / vs->ppt + 255) / 256)), dim3(256), kargs, 0, ctx->s_cmp);
      else
        e = cudaLaunchKernel((const void*)v->kernel, dim3(unsigned((cnt + 255) / 256)), dim3(256), kargs, 0, ctx->s_cmp);
      if (e != cudaSuccess) return jit_fail(m, EO_ERR_CUDA, "eo_jit_eval: launch: %s", cudaGetErrorString(e));
      ++ctx->launches;
    }
    return EO_OK;
  };
  return eo_run_streamed(ctx, args, na, n, launch);
}

// Fused evaluation: every operand is tabulated inside the kernel from the DOF coefficients of its eo_tab
// (kind per operand), so no operand array is ever written to or read from HBM.
int eo_jit_eval_tabulated(eo_jit* m, const int* derivatives, const double* params, eo_tab* const* tabs, const int* kinds,
                          const double* const* coefficients, const double* const* state, double* out, double* value,
                          double* const* aux) {
  if (!m) return EO_ERR_INVALID;
  eo_ctx* ctx = m->ctx;
  if (!ctx) return jit_fail(m, EO_ERR_NO_DEVICE, "eo_jit_eval_tabulated: this model was created without a context (compile-only)");
  if (!out || !tabs || !kinds || !coefficients || (m->n_state && !state) || (m->n_params && !params))
    return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval_tabulated: NULL argument");
  int order, da, db;
  int rc = jit_multi_index(m, derivatives, order, da, db);
  if (rc) return rc;
  EO_CUDA(ctx, cudaSetDevice(ctx->device));
  eo_tab_view tv[EO_JIT_MAX_ARGS];
  jit_tab_sig sig[EO_JIT_MAX_ARGS];
  int nq = 0;
  int64_t n_cells = 0;
  for (int i = 0; i < m->n_operands; ++i) {
    if (!tabs[i] || !coefficients[i]) return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval_tabulated: operand %d has no tabulation source", i);
    const int nc = eo_tab_ncomp(tabs[i], kinds[i]);
    if (nc != m->operand_size[i])
      return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval_tabulated: operand %d has %d components, its tabulation kind gives %d", i,
                      m->operand_size[i], nc);
    rc = eo_tab_view_get(tabs[i], coefficients[i], &tv[i], i);
    if (rc) return rc;
    if (tv[i].ctx != ctx) return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval_tabulated: operand %d lives on another context", i);
    const tab_tables* T = tv[i].T_host;
    sig[i] = {T->gdim, T->bs, T->nb, kinds[i]};
    if (i == 0) nq = T->nq, n_cells = tv[i].n_cells;
    if (T->nq != nq || tv[i].n_cells != n_cells)
      return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval_tabulated: all operands must share the cells and the evaluation points");
  }
  jit_variant* v = nullptr;
  rc = jit_compile(m, order, da, db, &v, nq, sig);
  if (rc) return rc;
  const int64_t n = n_cells * nq;

  eo_arg args[2 * EO_JIT_MAX_ARGS + 2];
  int na = 0;
  for (int i = 0; i < m->n_state; ++i) {
    if (!state[i]) return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval_tabulated: state field %d is NULL", i);
    args[na++] = {state[i], size_t(m->state_size[i]) * 8, false};
  }
  const int i_out = na;
  args[na++] = {out, size_t(v->out_width) * 8, true};
  const int i_val = na;
  args[na++] = {order >= 1 ? value : nullptr, size_t(m->out_size) * 8, true};
  const int i_aux = na;
  for (int i = 0; i < m->n_aux; ++i) args[na++] = {aux ? aux[i] : nullptr, size_t(m->aux_size[i]) * 8, true};

  eo_jit_fused_args ka;
  memset(&ka, 0, sizeof ka);
  for (int i = 0; i < m->n_params; ++i) ka.prm[i] = params[i];
  for (int i = 0; i < m->n_operands; ++i) {
    ka.T[i] = tv[i].T_dev;
    ka.dofmap[i] = tv[i].dofmap, ka.x_dofmap[i] = tv[i].x_dofmap, ka.x[i] = tv[i].x, ka.u[i] = tv[i].u;
  }
  cudaKernel_t kernel = v->kernel;
  auto launch = [&](void** p, int64_t nn, int64_t off) -> int {
    for (int i = 0; i < na; ++i)
      if (p[i] && (reinterpret_cast<uintptr_t>(p[i]) & (args[i].bpq % 32 == 0 ? 31 : args[i].bpq % 16 == 0 ? 15 : 7)))
        return jit_fail(m, EO_ERR_INVALID, "eo_jit_eval_tabulated: device argument %d is misaligned for its %zu-byte points", i, args[i].bpq);
    for (int i = 0; i < m->n_state; ++i) ka.state[i] = (const double*)p[i];
    ka.out = (double*)p[i_out];
    ka.value = (double*)p[i_val];
    for (int i = 0; i < m->n_aux; ++i) ka.aux[i] = (double*)p[i_aux + i];
    ka.n = nn;
    ka.point_offset = off;
    void* kargs[1] = {&ka};
    cudaError_t e = cudaLaunchKernel((const void*)kernel, dim3(unsigned((nn + 255) / 256)), dim3(256), kargs, 0, ctx->s_cmp);
    if (e != cudaSuccess) return jit_fail(m, EO_ERR_CUDA, "eo_jit_eval_tabulated: launch: %s", cudaGetErrorString(e));
    ++ctx->launches;
    return EO_OK;
  };
  return eo_run_streamed(ctx, args, na, n, launch);
}

// compile-only counterpart (no GPU needed): element signature given explicitly
int eo_jit_compile_tabulated(eo_jit* m, const int* derivatives, int nq, const int* gdim, const int* bs, const int* nb,
                             const int* kinds, size_t* cubin_bytes) {
  if (!m || !gdim || !bs || !nb || !kinds) return EO_ERR_INVALID;
  int order, da, db;
  int rc = jit_multi_index(m, derivatives, order, da, db);
  if (rc) return rc;
  jit_tab_sig sig[EO_JIT_MAX_ARGS];
  for (int i = 0; i < m->n_operands; ++i) {
    sig[i] = {gdim[i], bs[i], nb[i], kinds[i]};
    if (!((gdim[i] == 2 || gdim[i] == 3) && bs[i] >= 1 && bs[i] <= EO_TAB_MAX_BS && nb[i] >= 1 && nb[i] <= EO_TAB_MAX_NB &&
          tab_ncomp(kinds[i], bs[i], gdim[i]) == m->operand_size[i]))
      return jit_fail(m, EO_ERR_INVALID, "eo_jit_compile_tabulated: bad element signature for operand %d", i);
  }
  if (nq < 1 || nq > EO_TAB_MAX_NQ) return jit_fail(m, EO_ERR_INVALID, "eo_jit_compile_tabulated: nq out of range");
  jit_variant* v = nullptr;
  rc = jit_compile(m, order, da, db, &v, nq, sig);
  if (rc) return rc;
  if (cubin_bytes) *cubin_bytes = v->cubin.size();
  return EO_OK;
}

}  // extern "C"
