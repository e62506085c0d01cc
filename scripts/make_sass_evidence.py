#!/usr/bin/env python
"""Static evidence for profiles/: per-kernel SASS instruction counts (tcgen05 / TMEM / TMA mnemonics, peer-store
variants) and the `-Xptxas -v` resource table, from the objects of the last build.  Runs on the CPU box
(cuobjdump / the ptxas logs), no GPU needed:   python scripts/make_sass_evidence.py"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "rxinfer.jl_b200", "build")
OUT = os.path.join(ROOT, "profiles")
MNEMONICS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "HMMA", "DFMA", "FFMA", "MUFU", "STG", "LDG", "BAR", "SYNCS")


def demangle(names):
    p = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return p.stdout.splitlines()


def main():
    rows = []
    for obj in sorted(glob.glob(os.path.join(BUILD, "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        cur, counts = None, {}
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                counts[cur] = {"n": 0}
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and cur:
                counts[cur]["n"] += 1
                op = m.group(1).split(".")[0]
                for mn in MNEMONICS:
                    if op == mn:
                        counts[cur][mn] = counts[cur].get(mn, 0) + 1
        for k, c in counts.items():
            rows.append((os.path.basename(obj), k, c))
    names = demangle([r[1] for r in rows])
    keep = re.compile(r"lgssm_shared_kernel<4, 4, 2, 4, true, false, false, true|lgssm_seg_kernel<4, 4|lgssm_umma_sweep|umma_ky_kernel|"
                      r"umma_selftest|hgf_filter_kernel|gain_scan_kernel<4, 4>|lgssm_chain_kernel<4, 4|lgssm_generic_chain_kernel|"
                      r"peer_|replicate_cov|large_gain_tables<64|large_fwd_doubling<64|broadcast_cov|mv_iid_wishart_vmp_kernel<2>|seg_tables_kernel<4, 4>|"
                      r"lar_vmp_kernel<[15]>|k_left_gemm<64>|k_cholinv_warp|k_ew")
    with open(os.path.join(OUT, "r2_sass_summary.txt"), "w") as f:
        f.write("# SASS instruction counts per kernel (cuobjdump -sass of rxinfer.jl_b200/build/*.o, sm_100a); made by scripts/make_sass_evidence.py\n")
        f.write("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk), LDGSTS = cp.async\n")
        for (obj, _, c), name in zip(rows, names):
            if not keep.search(name):
                continue
            short = re.sub(r"\(.*", "", name)
            extra = " ".join(f"{k}={v}" for k, v in c.items() if k != "n" and v)
            f.write(f"{obj:28s} {c['n']:6d} instr  {short}\n{'':28s}        {extra}\n")
    with open(os.path.join(OUT, "r2_ptxas_table.txt"), "w") as f:
        f.write("# registers / shared memory / spills per kernel (-Xptxas -v logs of the build); made by scripts/make_sass_evidence.py\n")
        for log in sorted(glob.glob(os.path.join(BUILD, "*.ptxas.log"))):
            txt = open(log).read()
            ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                              r"ptxas info\s+: Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", txt)
            dn = demangle([e[0] for e in ents])
            for e, name in zip(ents, dn):
                if not keep.search(name):
                    continue
                short = re.sub(r"\(.*", "", name)
                f.write(f"{os.path.basename(log):32s} regs={e[4]:>3s} smem={e[6] or '0':>6s} stack={e[1]:>5s} spill_st={e[2]} spill_ld={e[3]}  {short}\n")
    print(open(os.path.join(OUT, "r2_sass_summary.txt")).read()[:3000])


if __name__ == "__main__":
    main()
