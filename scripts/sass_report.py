"""Per-kernel SASS evidence from the shipped library (no GPU needed): for every kernel in `_C/librealhf_b200_ops.so` a mnemonic
histogram and the lines that carry the Blackwell-specific instructions (tcgen05 -> UTC*MMA / UTCBAR / LDTM, TMA -> UTMALDG,
NVSwitch multicast -> LDGMC / STGMC / REDGMC ...).  Writes `profiles/sass/<family>.txt` + `profiles/sass/summary.json`.

    python scripts/sass_report.py
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "realhf_b200", "_C", "librealhf_b200_ops.so")
OUT = os.path.join(ROOT, "profiles", "sass")
KEY = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|UTCBAR[.\w]*|UTCCP[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|LDGMC[.\w]*|STGMC[.\w]*|REDGMC[.\w]*|"
                 r"MULTIMEM[.\w]*|SYNCS[.\w]*|ACQBULK|CCTL[.\w]*|HMMA[.\w]*|ERRBAR|MEMBAR[.\w]*)")
FAMILIES = [("gemm_2cta", r"gemm_2cta_kernel"), ("gemm_tile", r"gemm_kernel|gemm_tcgen05_kernel"), ("gemm_streamk", r"gemm_streamk_kernel"),
            ("gemm_grouped", r"grouped"), ("attn_fwd", r"attn_fwd_kernel"), ("attn_bwd", r"attn_bwd_kernel|attn_delta"), ("attn_decode", r"decode_attn"),
            ("nvls", r"nvls_"), ("allreduce", r"allreduce_|barrier_kernel|reduce_slabs|spin_wait"), ("ep", r"ep_plan|ep_move"),
            ("segcopy", r"segcopy|segment_copy"), ("adam", r"adamw_kernel|sumsq"), ("norm", r"rmsnorm|layernorm|colsum"),
            ("sampling", r"sample_kernel"), ("logprob", r"logprob"), ("gae", r"gae"), ("elementwise", r"rope_kernel|gated_act")]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    os.makedirs(OUT, exist_ok=True)
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
        elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            kernels[cur].append(line)
    demangled = {}
    names = list(kernels)
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, out):
        demangled[n] = d
    summary = {}
    for fam, pat in FAMILIES:
        rx = re.compile(pat)
        chosen = [n for n in names if rx.search(demangled.get(n, n))]
        if not chosen:
            continue
        with open(os.path.join(OUT, f"{fam}.txt"), "w") as f:
            f.write(f"# SASS evidence for `{fam}` kernels of {os.path.basename(LIB)} (cuobjdump -sass, sm_100a)\n")
            for n in chosen:
                lines = kernels[n]
                hist = collections.Counter()
                key_lines = []
                for ln in lines:
                    body = re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip()
                    op = body.split(";")[0].split()
                    if not op:
                        continue
                    mn = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
                    hist[mn.split(".")[0]] += 1
                    if KEY.search(body):
                        key_lines.append(ln.rstrip())
                keyhist = collections.Counter()
                for ln in key_lines:
                    keyhist[KEY.search(ln).group(1)] += 1
                short = re.sub(r"\(anonymous namespace\)::", "", demangled[n])[:200]
                f.write(f"\n## {short}\n   {len(lines)} instructions; top mnemonics: " + ", ".join(f"{k} x{v}" for k, v in hist.most_common(12)) + "\n")
                f.write("   Blackwell / NVSwitch instructions: " + (", ".join(f"{k} x{v}" for k, v in sorted(keyhist.items())) or "none") + "\n")
                for ln in key_lines[:24]:
                    f.write(ln + "\n")
                if len(key_lines) > 24:
                    f.write(f"        ... {len(key_lines) - 24} more\n")
                summary.setdefault(fam, {})[short[:120]] = dict(n_instr=len(lines), key=dict(keyhist))
    with open(os.path.join(OUT, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    tot = collections.Counter()
    for fam in summary.values():
        for k in fam.values():
            for mn, c in k["key"].items():
                tot[mn.split(".")[0]] += c
    print(json.dumps(dict(kernels=sum(len(v) for v in summary.values()), families=len(summary), totals=dict(tot))))


if __name__ == "__main__":
    sys.exit(main())
