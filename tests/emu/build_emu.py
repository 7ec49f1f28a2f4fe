"""TEST INFRASTRUCTURE: build tests/emu/_build/libkge_emu.so — the kernels of mkb_b200/csrc compiled
for the HOST against the CUDA execution-model emulation in tests/emu/include/cuda_runtime.h.

The product sources are not modified: this script rewrites, on the fly, the three constructs g++
cannot parse

  * ``kernel<<<grid, block, smem, stream>>>(args)``  ->  ``emu::launch(&kernel, grid, block, smem, stream, [&]{ kernel(args); })``
  * ``extern __shared__ [__align__(n)] T name[];``   ->  ``T* name = (T*)emu::S.dyn_smem;``
  * the inline-PTX statements (red.*.add[.v4].f32, sqrt/rsqrt.approx) -> plain C++

and compiles everything except rank_tc.cu (tcgen05/TMA PTX cannot be emulated; kge_rank_all then
falls back to the fp32 tile kernel, which is what the emulation tests cover).

Memory-safety check of every kernel (out-of-bounds reads and writes on the "device" buffers):

    python tests/emu/build_emu.py --asan          # -> tests/emu/_build_asan/libkge_emu.so
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
    ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1:verify_asan_link_order=0 \
    KGE_EMU_LIB=tests/emu/_build_asan/libkge_emu.so python -m pytest tests/test_emu_kernels.py -q
"""
import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "mkb_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libkge_emu.so")
SOURCES = ["api.cu", "loss.cu", "sampler.cu", "score.cu", "rank.cu", "topk.cu", "byent.cu", "pooled.cu", "peer.cu"]
HEADERS = ["kge_common.cuh"]


def _balanced(text, start, open_ch="(", close_ch=")"):
    """Index just past the bracket that closes the one at text[start]."""
    assert text[start] == open_ch, text[start:start + 20]
    depth, i, in_str = 0, start, False
    while i < len(text):
        c = text[i]
        if in_str:
            if c == "\\":
                i += 1
            elif c == '"':
                in_str = False
        elif c == '"':
            in_str = True
        elif c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced brackets")


def rewrite_launches(text):
    out, pos = [], 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            out.append(text[pos:])
            return "".join(out)
        # kernel expression: identifier, optionally followed by a <template, argument, list>
        j = k
        while j > pos and text[j - 1].isspace():
            j -= 1
        if text[j - 1] == ">":
            depth, j = 0, j - 1
            while True:
                if text[j] == ">":
                    depth += 1
                elif text[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        while j > pos and (text[j - 1].isalnum() or text[j - 1] in "_:"):
            j -= 1
        kernel = text[j:k].strip()
        e = text.find(">>>", k)
        cfg = text[k + 3:e]
        a0 = e + 3
        while text[a0].isspace():
            a0 += 1
        a1 = _balanced(text, a0)
        args = text[a0 + 1:a1 - 1]
        parts = [c.strip() for c in _split_top(cfg)]
        while len(parts) < 4:
            parts.append("0")
        out.append(text[pos:j])
        out.append(f"emu::launch(emu::fn_addr({kernel}), {parts[0]}, {parts[1]}, (size_t)({parts[2]}), "
                   f"(cudaStream_t)({parts[3]}), [&]() {{ {kernel}({args}); }})")
        pos = a1


def _split_top(s):
    parts, depth, cur = [], 0, []
    for c in s:  # launch configurations hold casts and member accesses (`t->n`), never template arguments
        if c in "([":
            depth += 1
        elif c in ")]":
            depth -= 1
        if c == "," and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(c)
    parts.append("".join(cur))
    return parts


def rewrite_extern_shared(text):
    return re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];",
                  r"\1* \2 = reinterpret_cast<\1*>(emu::S.dyn_smem);", text)


def rewrite_static_shared(text):
    """``__shared__ T a[n], b;`` -> the same + ``emu::poison_shared(&a, sizeof(a)); ...``: a block starts with
    garbage in shared memory, not with what the previous block left there."""
    def fix(m):
        decl = m.group(0)
        names = []
        for d in _split_top(m.group(2)):
            names.append(re.match(r"\s*(\w+)", d).group(1))
        return decl + "".join(f" emu::poison_shared(&{n}, sizeof({n}));" for n in names)

    return re.sub(r"(?m)^(\s*__shared__\s+(?:__align__\(\d+\)\s+)?(?:unsigned\s+long\s+long|unsigned\s+int|\w+)\s+)"
                  r"([^;()]+);", fix, text)


def rewrite_asm(text):
    out, pos = [], 0
    for m in re.finditer(r"\basm\s*(?:volatile\s*)?\(", text):
        if m.start() < pos:
            continue
        end = _balanced(text, m.end() - 1)
        body = text[m.end():end - 1]
        semi = text.index(";", end)
        lit = re.match(r'\s*"((?:[^"\\]|\\.)*)"', body)
        ptx, rest = lit.group(1), body[lit.end():]
        ops, i = [], 0
        while True:  # operands: "constraint"(expression) with balanced parentheses
            mm = re.compile(r'"(=?[a-z]+)"\s*\(').search(rest, i)
            if not mm:
                break
            if mm.group(1) == "memory":
                i = mm.end()
                continue
            j = _balanced(rest, mm.end() - 1)
            ops.append((mm.group(1), rest[mm.end():j - 1]))
            i = j
        outs = [e for c, e in ops if c.startswith("=")]
        ins = [e for c, e in ops if not c.startswith("=")]
        if re.match(r"red\.relaxed\.(gpu|sys)\.global\.add\.v4\.f32", ptx):
            p, v = ins[0], ins[1:5]
            code = "{ float* _p = (float*)(" + p + "); if ((uintptr_t)_p & 15) { fprintf(stderr, \"cuda_emu: misaligned " \
                   "red.v4.f32\\n\"); abort(); } " + " ".join(f"_p[{i}] += ({e});" for i, e in enumerate(v)) + " }"
        elif re.match(r"red\.relaxed\.(gpu|sys)\.global\.add\.f32", ptx):
            code = f"{{ *(float*)({ins[0]}) += ({ins[1]}); }}"
        elif ptx.startswith("sqrt.approx"):
            code = f"{outs[0]} = sqrtf({ins[0]});"
        elif ptx.startswith("rsqrt.approx"):
            code = f"{outs[0]} = 1.0f / sqrtf({ins[0]});"
        elif re.match(r"(bar\.sync \d|fence\.proxy\.async|cp\.async\.bulk|cp\.reduce\.async\.bulk)", ptx):
            # K3b's staged bulk reductions (score.cu, opt-in with KGE_BWD_BULK=1) have no host model
            code = '{ fprintf(stderr, "cuda_emu: TMA bulk / named-barrier PTX is not emulated (KGE_BWD_BULK)\\n"); abort(); }'
        elif "globaltimer" in ptx:  # nanosecond wall clock
            code = (f"{{ struct timespec _ts; clock_gettime(CLOCK_MONOTONIC, &_ts); "
                    f"{outs[0]} = (long long)_ts.tv_sec * 1000000000LL + _ts.tv_nsec; }}")
        else:
            raise ValueError(f"no emulation for PTX statement: {ptx}")
        out.append(text[pos:m.start()])
        out.append(code)
        pos = semi + 1
    out.append(text[pos:])
    return "".join(out)


def preprocess(name):
    text = open(os.path.join(CSRC, name)).read()
    text = text.replace('#include "../../include/kge_b200.h"', f'#include "{os.path.join(ROOT, "include", "kge_b200.h")}"')
    text = rewrite_asm(rewrite_static_shared(rewrite_extern_shared(rewrite_launches(text))))
    if name == "score.cu":
        text = "#define KGE_NO_TMA 1  // score_tma.cuh (cp.async.bulk / mbarrier) is not emulated\n" + text
    return f"// GENERATED by tests/emu/build_emu.py from mkb_b200/csrc/{name} — do not edit\n" + text


STUB = """// rank_tc.cu (tcgen05 / TMA) is not emulated: report "unsupported" so kge_rank_all uses the tile kernel
#include "kge_common.cuh"
namespace kge {
int rank_tc_launch(const float*, const float*, int64_t, int, const int64_t*, int, const kge_filter_csr_t*, bool,
                   float*, const int64_t*, unsigned long long*, float*, bool, cudaStream_t, const float*, int) {
  return KGE_E_UNSUPPORTED;
}
bool rank_tc_eligible(const float*, int, int64_t) { return false; }
}  // namespace kge
extern "C" long kge_emu_launch_count(void) { return emu::S.launches; }
"""


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps += [os.path.join(ROOT, "include", "kge_b200.h"), os.path.join(HERE, "include", "cuda_runtime.h"), __file__]
    return any(os.path.getmtime(f) > t for f in deps)


def build(force=False, asan=False):
    global OUT, LIB
    if asan:
        OUT = os.path.join(HERE, "_build_asan")
        LIB = os.path.join(OUT, "libkge_emu.so")
        force = True
    if not force and not stale():
        return LIB
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    for h in HEADERS:
        open(os.path.join(OUT, h), "w").write(preprocess(h))
    cpps = []
    for s in SOURCES:
        dst = os.path.join(OUT, s[:-3] + ".cpp")
        open(dst, "w").write(preprocess(s))
        cpps.append(dst)
    stub = os.path.join(OUT, "rank_tc_stub.cpp")
    open(stub, "w").write(STUB)
    cpps.append(stub)
    flags = ["-std=c++17", "-O1", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-w",
             "-I", os.path.join(HERE, "include"), "-I", OUT]
    if asan:
        # + UBSan: misaligned vector / 64-bit accesses fault on the device, signed overflow is an indexing bug
        flags += ["-g", "-fsanitize=address,alignment,signed-integer-overflow,shift,bounds",
                  "-fno-sanitize-recover=all", "-fno-omit-frame-pointer"]
    procs = []
    for c in cpps:
        o = c[:-4] + ".o"
        procs.append((c, o, subprocess.Popen(["g++", *flags, "-c", c, "-o", o], stdout=subprocess.PIPE,
                                             stderr=subprocess.STDOUT, text=True)))
    objs = []
    for c, o, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"g++ failed on {c}:\n{log[-4000:]}")
        objs.append(o)
    subprocess.run(["g++", "-shared", *(["-fsanitize=address,undefined"] if asan else []), "-o", LIB, *objs], check=True)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, asan="--asan" in sys.argv))
