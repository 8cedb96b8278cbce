// jit.cu — run-time specialisation of the fused evaluator (host code).
//
// rustc monomorphises every `Zip<Map<..>>` type into its own collect loop.  The library ships that
// specialisation pre-built for the common op trees (sigs.hpp); for every other chain it builds it on
// first use: the op sequence of the device program becomes a compile-time signature, NVRTC instantiates
// the SAME hand-written evaluator (exec.cuh, embedded in the library at build time) for it —
// `eval_vector<JitSig, S, V, MAXD, WIDE, MAXR>` with the exact slot width, vector width, stack depth and
// rank of the plan — for sm_100a, and the cubin is cached per process.  Without NVRTC (or if compilation
// fails) the depth-specialised interpreter runs instead; nothing ever falls back to the CPU.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "program.hpp"

namespace mdim {

namespace {

#include "build/jit_sources.inc"  // kSrcExec, kSrcProgram, kSrcMdimH: the sources as raw string literals

const char kStdint[] =
    "#pragma once\n"
    "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
    "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n"
    "typedef unsigned long long uintptr_t;\n"
    "#define INT32_MIN (-2147483647 - 1)\n#define INT32_MAX 2147483647\n#define UINT32_MAX 4294967295u\n"
    "#define INT64_MIN (-9223372036854775807ll - 1)\n#define INT64_MAX 9223372036854775807ll\n#define UINT64_MAX 18446744073709551615ull\n";
const char kStddef[] = "#pragma once\ntypedef decltype(sizeof(0)) size_t;\n";
const char kString[] = "#pragma once\n";  // memcpy is a device builtin

struct Nvrtc {
    void* lib = nullptr;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    decltype(&nvrtcVersion) version = nullptr;
    bool ok = false;
    bool has_vec256 = false;  // PTX ISA 8.8 (CUDA 12.9): 256-bit vector loads / stores
    Nvrtc() {
        // the toolkit's own copy first: a process that imported torch already has torch's bundled (older) NVRTC
        // loaded under the same soname, and a lookup by name would hand that one back
        const char* names[] = {"/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so.12", "libnvrtc.so"};
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (lib) break; }
        if (!lib) return;
        create = (decltype(create))dlsym(lib, "nvrtcCreateProgram");
        compile = (decltype(compile))dlsym(lib, "nvrtcCompileProgram");
        destroy = (decltype(destroy))dlsym(lib, "nvrtcDestroyProgram");
        cubin_size = (decltype(cubin_size))dlsym(lib, "nvrtcGetCUBINSize");
        cubin = (decltype(cubin))dlsym(lib, "nvrtcGetCUBIN");
        log_size = (decltype(log_size))dlsym(lib, "nvrtcGetProgramLogSize");
        log = (decltype(log))dlsym(lib, "nvrtcGetProgramLog");
        version = (decltype(version))dlsym(lib, "nvrtcVersion");
        ok = create && compile && destroy && cubin_size && cubin && log_size && log;
        int major = 0, minor = 0;
        if (version && version(&major, &minor) == NVRTC_SUCCESS) has_vec256 = major > 12 || (major == 12 && minor >= 9);
    }
};

struct Entry { cudaKernel_t kernel = nullptr; cudaLibrary_t library = nullptr; bool failed = false; };

std::mutex g_mu;
std::map<std::string, Entry>& cache() { static std::map<std::string, Entry> c; return c; }
bool verbose() { static const bool v = [] { const char* e = getenv("MDIM_JIT_VERBOSE"); return e && e[0] == '1'; }(); return v; }

// The shape-like part of the program as a constant-initialised device object (see MDIM_SHAPE_OF in exec.cuh).
std::string make_shape_source(const Program& P) {
    std::string s = "#include \"program.hpp\"\nnamespace mdim { __host__ __device__ constexpr Program mdim_make_shape() { Program p{};\n";
    char b[320];
    auto set = [&](const char* fmt, auto... args) { snprintf(b, sizeof b, fmt, args...); s += b; };
    set("p.rank=%d; p.red_rank=%d; p.n_instr=%d; p.n_addr=%d; p.n_pred=%d; p.out_dtype=%d; p.vec=%d; p.vpt=%d;\n", P.rank, P.red_rank, P.n_instr, P.n_addr,
        P.n_pred, P.out_dtype, P.vec, P.vpt);
    set("p.n_out=%d; p.out_dtypes[0]=%d; p.out_dtypes[1]=%d; p.out_dtypes[2]=%d; p.out_dtypes[3]=%d;\n", P.n_out, P.out_dtypes[0], P.out_dtypes[1], P.out_dtypes[2], P.out_dtypes[3]);
    set("p.n_vec=%lluull; p.red_count=%lluull; p.red_fast_len=%lluull;\n", (unsigned long long)P.n_vec, (unsigned long long)P.red_count,
        (unsigned long long)P.red_fast_len);
    for (int a = 0; a < kMaxRank; ++a)
        set("p.length[%d]=%lluull; p.dec_len[%d]=%lluull; p.dec_scale[%d]=%uu; p.div_mul[%d]=%uu; p.div_shr[%d]=%uu;\n", a, (unsigned long long)P.length[a], a,
            (unsigned long long)P.dec_len[a], a, P.dec_scale[a], a, P.div_mul[a], a, P.div_shr[a]);
    for (int i = 0; i < P.n_instr; ++i) {
        const Instr& I = P.instr[i];
        set("p.instr[%d].opc=%d; p.instr[%d].dtype=%d; p.instr[%d].op=%d; p.instr[%d].aux=%d; p.instr[%d].slot=%d; p.instr[%d].n=%d;\n", i, (int)I.opc, i, (int)I.dtype, i,
            (int)I.op, i, (int)I.aux, i, (int)I.slot, i, (int)I.n);
    }
    for (int i = 0; i < P.n_addr; ++i) {
        const Addr& A = P.addr[i];
        set("p.addr[%d].inner=%lldll; p.addr[%d].rstride=%lldll; p.addr[%d].n_peers=%d;\n", i, (long long)A.inner, i, (long long)A.rstride, i, A.n_peers);
        for (int a = 0; a < kMaxRank; ++a) if (A.stride[a]) set("p.addr[%d].stride[%d]=%lldll;", i, a, (long long)A.stride[a]);
        for (int c = 0; c < kMaxComp; ++c) if (A.gstride[c] || A.bound[c]) set("p.addr[%d].gstride[%d]=%lldll; p.addr[%d].bound[%d]=%lluull;", i, c, (long long)A.gstride[c], i, c, (unsigned long long)A.bound[c]);
        s += "\n";
    }
    for (int i = 0; i < P.n_pred; ++i) {
        const Pred& Q = P.pred[i];
        set("p.pred[%d].lane_coef=%d; p.pred[%d].cmp=%d; p.pred[%d].rcoef=%d; p.pred[%d].rhs=%lldll;", i, Q.lane_coef, i, Q.cmp, i, Q.rcoef, i, (long long)Q.rhs);
        for (int a = 0; a < kMaxRank; ++a) if (Q.coef[a]) set("p.pred[%d].coef[%d]=%d;", i, a, Q.coef[a]);
        s += "\n";
    }
    s += "return p; } }\nstatic __device__ const mdim::Program mdim_jit_shape = mdim::mdim_make_shape();\n#define MDIM_SHAPE_OF(P) (mdim_jit_shape)\n";
    return s;
}

std::string make_source(const Plan& p, int maxr, int maxd, bool with_shape) {
    std::string s = with_shape ? make_shape_source(p.prog) : std::string();
    if (const char* e = getenv("MDIM_JIT_STORE_DEFAULT")) if (e[0] == '1') s = "#define MDIM_STORE_STREAMING false\n" + s;
    s += "#include \"exec.cuh\"\nnamespace mdim { struct JitSig { static constexpr SigInstr code[] = {";
    char buf[64];
    for (int i = 0; i < p.prog.n_instr; ++i) {
        const Instr& I = p.prog.instr[i];
        snprintf(buf, sizeof buf, "{%d,%d,%d,%d},", (int)I.opc, (int)I.dtype, (int)I.op, (int)I.aux);
        s += buf;
    }
    snprintf(buf, sizeof buf, "}; static constexpr int n = %d; }; }\n", p.prog.n_instr);
    s += buf;
    char k[1200];
    snprintf(k, sizeof k,
             "extern \"C\" __global__ void __launch_bounds__(256) mdim_jit_kernel(const __grid_constant__ mdim::Program P, void* __restrict__ out, "
             "mdim::ErrWord* __restrict__ err, unsigned long long g_begin, unsigned long long g_end) {\n"
             "  if (!(P.flags & mdim::PF_NOWAIT) || (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)) asm volatile(\"griddepcontrol.wait;\" ::: \"memory\");\n"
             "  asm volatile(\"griddepcontrol.launch_dependents;\" ::: \"memory\");\n"
             "  const unsigned long long step = (unsigned long long)gridDim.x * 256ull;\n"
             "  for (unsigned long long g = g_begin + (unsigned long long)blockIdx.x * 256ull + threadIdx.x; g < g_end; g += step)\n"
             "    mdim::eval_vector<mdim::JitSig, %s, %d, %d, %s, %d, 1>(P, out, err, g);\n}\n",
             p.slot_bytes == 8 ? "uint64_t" : "uint32_t", p.vec, maxd, p.wide ? "true" : "false", maxr);
    s += k;
    return s;
}

// cache key of the shape-like part: the program with every run-time field (pointers, offsets, immediates) zeroed
std::string shape_key(const Program& P) {
    Program q = P;
    q.flags &= PF_VEC256; q.explain_pos = 0;
    for (int i = 0; i < kMaxInstr; ++i) q.instr[i].imm = 0;
    for (int i = 0; i < kMaxAddr; ++i) { q.addr[i].ptr = nullptr; q.addr[i].offset = 0; }
    memset(&q.peers, 0, sizeof q.peers);
    memset(q.out_more, 0, sizeof q.out_more);
    return std::string((const char*)&q, sizeof q);
}

}  // namespace

// ---- on-disk cache of compiled cubins -------------------------------------------------------------------------
// A chain costs ~0.2 s of NVRTC the first time a PROCESS sees it; the cubin only depends on the generated source,
// the embedded headers, the options and the compiler version, so when MDIM_JIT_CACHE names a directory it is
// kept there as <fnv1a-64 of all of those>.cubin (opt-in: the library writes nowhere it was not told to).
static uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
static std::string cache_dir() {
    static const std::string dir = [] {
        const char* e = getenv("MDIM_JIT_CACHE");
        if (!e || !e[0] || (e[0] == '0' && !e[1])) return std::string();
        const std::string d = e;
        mkdir(d.c_str(), 0700);
        struct stat st;
        return stat(d.c_str(), &st) == 0 && S_ISDIR(st.st_mode) && access(d.c_str(), W_OK) == 0 ? d : std::string();
    }();
    return dir;
}
static bool cache_read(const std::string& path, std::vector<char>& cubin) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    bool ok = n > 64;
    if (ok) { cubin.resize((size_t)n); ok = fread(cubin.data(), 1, (size_t)n, f) == (size_t)n && memcmp(cubin.data(), "\x7f" "ELF", 4) == 0; }
    fclose(f);
    if (!ok) cubin.clear();
    return ok;
}
static void cache_write(const std::string& path, const std::vector<char>& cubin) {
    char tmp[600];
    snprintf(tmp, sizeof tmp, "%s.%d.tmp", path.c_str(), (int)getpid());
    FILE* f = fopen(tmp, "wb");
    if (!f) return;
    const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    fclose(f);
    if (!ok || rename(tmp, path.c_str()) != 0) unlink(tmp);  // rename is atomic: readers never see a partial file
}

// NVRTC: source -> sm_100a cubin.  Returns false (and the compiler log) on failure.
static bool compile_cubin(const Plan& p, int maxr, int maxd, bool with_shape, std::vector<char>& cubin, std::string& log_out) {
    static Nvrtc nv;
    if (!nv.ok) { log_out = "NVRTC (libnvrtc.so.12) is not available"; return false; }
    const std::string src = make_source(p, maxr, maxd, with_shape);
    const char* headers[] = {kSrcExec, kSrcProgram, kSrcMdimH, kStdint, kStddef, kString};
    const char* names[] = {"exec.cuh", "program.hpp", "../../include/mdim.h", "stdint.h", "stddef.h", "string.h"};
    nvrtcProgram prog = nullptr;
    if (nv.create(&prog, src.c_str(), "mdim_jit.cu", 6, headers, names) != NVRTC_SUCCESS) { log_out = "nvrtcCreateProgram failed"; return false; }
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "--fmad=false", "-lineinfo", "-DMDIM_NO_VEC256"};
    const int n_opts = nv.has_vec256 ? 4 : 5;
    std::string cached;
    if (!cache_dir().empty() && !getenv("MDIM_JIT_DUMP")) {
        uint64_t h = 14695981039346656037ull;
        h = fnv1a(h, src.data(), src.size());
        for (const char* hd : headers) h = fnv1a(h, hd, strlen(hd));
        for (int i = 0; i < n_opts; ++i) h = fnv1a(h, opts[i], strlen(opts[i]) + 1);
        int ver[2] = {0, 0};
        if (nv.version) nv.version(&ver[0], &ver[1]);
        h = fnv1a(h, ver, sizeof ver);
        char name[64];
        snprintf(name, sizeof name, "/%016llx.cubin", (unsigned long long)h);
        cached = cache_dir() + name;
        if (cache_read(cached, cubin)) { nv.destroy(&prog); return true; }
    }
    const nvrtcResult rc = nv.compile(prog, n_opts, opts);
    if (rc != NVRTC_SUCCESS) {
        size_t n = 0; nv.log_size(prog, &n);
        std::vector<char> log(n + 1, 0); nv.log(prog, log.data());
        log_out = log.data();
        nv.destroy(&prog);
        return false;
    }
    size_t n = 0;
    if (nv.cubin_size(prog, &n) == NVRTC_SUCCESS && n) { cubin.resize(n); if (nv.cubin(prog, cubin.data()) != NVRTC_SUCCESS) cubin.clear(); }
    nv.destroy(&prog);
    if (cubin.empty()) { log_out = "no cubin produced"; return false; }
    if (!cached.empty()) cache_write(cached, cubin);
    if (const char* dir = getenv("MDIM_JIT_DUMP")) {  // for cuobjdump -sass
        static int counter = 0;
        char path[512];
        snprintf(path, sizeof path, "%s/mdim_jit_%d%s.cubin", dir, counter++, with_shape ? "_shape" : "");
        if (FILE* f = fopen(path, "wb")) { fwrite(cubin.data(), 1, cubin.size(), f); fclose(f); }
    }
    return true;
}

static void jit_shape(const Plan& p, int& maxr, int& maxd) {
    maxr = p.kind == KK_STREAM ? 1 : (p.n_axes < 1 ? 1 : p.n_axes);
    maxd = p.max_depth < 1 ? 1 : p.max_depth;
}

int jit_compile_check(const Plan& p, char* log, size_t log_len) {
    if ((p.kind != KK_STREAM && p.kind != KK_GENERIC) || p.vpt != 1) { if (log && log_len) snprintf(log, log_len, "not an evaluator plan"); return MDIM_ERR_UNSUPPORTED; }
    int maxr, maxd; jit_shape(p, maxr, maxd);
    std::vector<char> cubin; std::string msg;
    bool ok = compile_cubin(p, maxr, maxd, false, cubin, msg);
    if (ok && !p.wide) ok = compile_cubin(p, maxr, maxd, true, cubin, msg);  // and the shape-specialised form
    if (log && log_len) snprintf(log, log_len, "%s", ok ? "ok" : msg.c_str());
    if (ok) return MDIM_OK;
    return msg.rfind("NVRTC", 0) == 0 ? MDIM_ERR_UNSUPPORTED : MDIM_ERR_INVALID;
}

// Returns a kernel specialised for the plan, or nullptr (use the pre-built signature / the interpreter).
//   level 1: the op sequence (only for plans without a pre-built signature)
//   level 2: the op sequence AND the shape — strides, lengths, dividers and predicates become immediates.
//            Built when the same (ops, shape) has been collected MDIM_JIT_SHAPES times (default 2; 0 = never,
//            1 = at once) and the chain is a rank >= 2 evaluator plan, where the coordinate decode dominates.
void* jit_kernel_for(const Plan& p, int* level) {
    if (level) *level = 0;
    static const bool enabled = [] { const char* e = getenv("MDIM_JIT"); return !(e && e[0] == '0'); }();
    static const int shape_after = [] { const char* e = getenv("MDIM_JIT_SHAPES"); return e ? atoi(e) : 2; }();
    if (!enabled || (p.kind != KK_STREAM && p.kind != KK_GENERIC) || p.vpt != 1) return nullptr;
    int maxr, maxd; jit_shape(p, maxr, maxd);
    std::string key(p.sig, p.sig + p.sig_len);
    char tail[64];
    snprintf(tail, sizeof tail, "|%d|%d|%d|%d|%d", p.slot_bytes, p.vec, maxd, p.wide, maxr);
    key += tail;
    std::lock_guard<std::mutex> lock(g_mu);
    auto build = [&](Entry& e, bool with_shape) -> void* {
        if (e.kernel) return (void*)e.kernel;
        if (e.failed) return nullptr;
        e.failed = true;  // until proven otherwise
        std::vector<char> cubin; std::string msg;
        if (!compile_cubin(p, maxr, maxd, with_shape, cubin, msg)) {
            if (verbose()) fprintf(stderr, "mdim jit: %s — not specialised\n", msg.c_str());
            return nullptr;
        }
        cudaLibrary_t lib = nullptr;
        cudaKernel_t kern = nullptr;
        if (cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess ||
            cudaLibraryGetKernel(&kern, lib, "mdim_jit_kernel") != cudaSuccess) {
            cudaGetLastError();
            if (verbose()) fprintf(stderr, "mdim jit: loading the cubin failed — not specialised\n");
            return nullptr;
        }
        e.kernel = kern; e.library = lib; e.failed = false;
        if (verbose())
            fprintf(stderr, "mdim jit: specialised %d instructions (slot %d, V %d, depth %d, rank %d%s)%s\n", p.prog.n_instr, p.slot_bytes * 8, p.vec, maxd, maxr,
                    p.wide ? ", wide" : "", with_shape ? " for one shape" : "");
        return (void*)kern;
    };
    if (shape_after > 0 && p.kind == KK_GENERIC && !p.wide && !(p.prog.flags & PF_EXPLAIN)) {
        static std::map<std::string, int> seen;
        const std::string skey = key + "#" + shape_key(p.prog);
        if (seen.size() > 4096) seen.clear();  // a workload of ever-changing shapes must not grow this without bound
        if (++seen[skey] >= shape_after) {
            if (void* k = build(cache()[skey], true)) { if (level) *level = 2; return k; }
        }
    }
    if (p.static_id >= 0) return nullptr;  // the pre-built signature serves it
    void* k = build(cache()[key], false);
    if (k && level) *level = 1;
    return k;
}

}  // namespace mdim
