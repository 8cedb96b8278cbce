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
    bool ok = false;
    Nvrtc() {
        const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (lib) break; }
        if (!lib) return;
        create = (decltype(create))dlsym(lib, "nvrtcCreateProgram");
        compile = (decltype(compile))dlsym(lib, "nvrtcCompileProgram");
        destroy = (decltype(destroy))dlsym(lib, "nvrtcDestroyProgram");
        cubin_size = (decltype(cubin_size))dlsym(lib, "nvrtcGetCUBINSize");
        cubin = (decltype(cubin))dlsym(lib, "nvrtcGetCUBIN");
        log_size = (decltype(log_size))dlsym(lib, "nvrtcGetProgramLogSize");
        log = (decltype(log))dlsym(lib, "nvrtcGetProgramLog");
        ok = create && compile && destroy && cubin_size && cubin && log_size && log;
    }
};

struct Entry { cudaKernel_t kernel = nullptr; cudaLibrary_t library = nullptr; bool failed = false; };

std::mutex g_mu;
std::map<std::string, Entry>& cache() { static std::map<std::string, Entry> c; return c; }
bool verbose() { static const bool v = [] { const char* e = getenv("MDIM_JIT_VERBOSE"); return e && e[0] == '1'; }(); return v; }

std::string make_source(const Plan& p, int maxr, int maxd) {
    std::string s = "#include \"exec.cuh\"\nnamespace mdim { struct JitSig { static constexpr SigInstr code[] = {";
    char buf[64];
    for (int i = 0; i < p.prog.n_instr; ++i) {
        const Instr& I = p.prog.instr[i];
        snprintf(buf, sizeof buf, "{%d,%d,%d,%d},", (int)I.opc, (int)I.dtype, (int)I.op, (int)I.aux);
        s += buf;
    }
    snprintf(buf, sizeof buf, "}; static constexpr int n = %d; }; }\n", p.prog.n_instr);
    s += buf;
    char k[512];
    snprintf(k, sizeof k,
             "extern \"C\" __global__ void __launch_bounds__(256) mdim_jit_kernel(const __grid_constant__ mdim::Program P, void* __restrict__ out, "
             "mdim::ErrWord* __restrict__ err, unsigned long long g_begin, unsigned long long g_end) {\n"
             "  const unsigned long long step = (unsigned long long)gridDim.x * 256ull;\n"
             "  for (unsigned long long g = g_begin + (unsigned long long)blockIdx.x * 256ull + threadIdx.x; g < g_end; g += step)\n"
             "    mdim::eval_vector<mdim::JitSig, %s, %d, %d, %s, %d, 1>(P, out, err, g);\n}\n",
             p.slot_bytes == 8 ? "uint64_t" : "uint32_t", p.vec, maxd, p.wide ? "true" : "false", maxr);
    s += k;
    return s;
}

}  // namespace

// NVRTC: source -> sm_100a cubin.  Returns false (and the compiler log) on failure.
static bool compile_cubin(const Plan& p, int maxr, int maxd, std::vector<char>& cubin, std::string& log_out) {
    static Nvrtc nv;
    if (!nv.ok) { log_out = "NVRTC (libnvrtc.so.12) is not available"; return false; }
    const std::string src = make_source(p, maxr, maxd);
    const char* headers[] = {kSrcExec, kSrcProgram, kSrcMdimH, kStdint, kStddef, kString};
    const char* names[] = {"exec.cuh", "program.hpp", "../../include/mdim.h", "stdint.h", "stddef.h", "string.h"};
    nvrtcProgram prog = nullptr;
    if (nv.create(&prog, src.c_str(), "mdim_jit.cu", 6, headers, names) != NVRTC_SUCCESS) { log_out = "nvrtcCreateProgram failed"; return false; }
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "--fmad=false", "-lineinfo"};
    const nvrtcResult rc = nv.compile(prog, 4, opts);
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
    const bool ok = compile_cubin(p, maxr, maxd, cubin, msg);
    if (log && log_len) snprintf(log, log_len, "%s", ok ? "ok" : msg.c_str());
    if (ok) return MDIM_OK;
    return msg.rfind("NVRTC", 0) == 0 ? MDIM_ERR_UNSUPPORTED : MDIM_ERR_INVALID;
}

// Returns a kernel specialised for the plan's op sequence, or nullptr (use the interpreter).
void* jit_kernel_for(const Plan& p) {
    static const bool enabled = [] { const char* e = getenv("MDIM_JIT"); return !(e && e[0] == '0'); }();
    if (!enabled || (p.kind != KK_STREAM && p.kind != KK_GENERIC) || p.vpt != 1) return nullptr;
    int maxr, maxd; jit_shape(p, maxr, maxd);
    std::string key(p.sig, p.sig + p.sig_len);
    char tail[64];
    snprintf(tail, sizeof tail, "|%d|%d|%d|%d|%d", p.slot_bytes, p.vec, maxd, p.wide, maxr);
    key += tail;
    std::lock_guard<std::mutex> lock(g_mu);
    Entry& e = cache()[key];
    if (e.kernel) return (void*)e.kernel;
    if (e.failed) return nullptr;
    e.failed = true;  // until proven otherwise
    std::vector<char> cubin; std::string msg;
    if (!compile_cubin(p, maxr, maxd, cubin, msg)) {
        if (verbose()) fprintf(stderr, "mdim jit: %s — using the interpreter\n", msg.c_str());
        return nullptr;
    }
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kern = nullptr;
    if (cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess ||
        cudaLibraryGetKernel(&kern, lib, "mdim_jit_kernel") != cudaSuccess) {
        cudaGetLastError();
        if (verbose()) fprintf(stderr, "mdim jit: loading the cubin failed, using the interpreter\n");
        return nullptr;
    }
    e.kernel = kern; e.library = lib; e.failed = false;
    if (verbose()) fprintf(stderr, "mdim jit: specialised %d instructions (slot %d, V %d, depth %d, rank %d%s)\n", p.prog.n_instr, p.slot_bytes * 8, p.vec, maxd, maxr, p.wide ? ", wide" : "");
    return (void*)kern;
}

}  // namespace mdim
