"""Blackwell-native evidence from the built library: per kernel, how many tcgen05 / TMEM / TMA / mbarrier SASS instructions it
holds (`cuobjdump -sass trueno_b200/libtrueno_cuda.so`; the PTX names never appear in SASS — B200_PROFILING.md), plus the
first raw lines of each kind from the GEMM and attention kernels.
usage: python scripts/sass_excerpt.py > profiles/rNN_sass_excerpt.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "trueno_b200", "libtrueno_cuda.so")
KINDS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR.2CTA.MULTICAST", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCCP",
         "SYNCS.ARRIVE", "SYNCS.PHASECHK"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    fn, counts, samples = None, collections.OrderedDict(), collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            fn = re.sub(r"^void ", "", fn)
            counts.setdefault(fn, collections.Counter())
            continue
        if fn is None:
            continue
        body = re.sub(r"/\*[0-9a-f]+\*/", "", line).strip()
        for k in KINDS:
            if re.search(r"(^|\s)" + re.escape(k) + r"(\.|\s|$)", body):
                counts[fn][k] += 1
                if k in ("UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR.2CTA.MULTICAST") and \
                        any(t in fn for t in ("gemm_tf32x3_fused_astat", "attention_tf32x3_pair")):
                    samples.setdefault((fn, k), body.rstrip(";").strip())
                break
    print(f"# SASS excerpt of `trueno_b200/libtrueno_cuda.so` (cubins: {', '.join(arch)})\n")
    print("`tcgen05.mma` → `UTCHMMA` (`.2CTA` = `cta_group::2`), `tcgen05.commit` → `UTCBAR`, `tcgen05.ld/st` → `LDTM` / `STTM`, "
          "`cp.async.bulk.tensor` → `UTMALDG` / `UTMASTG`, `cp.async.bulk` → `UBLKCP`, mbarrier → `SYNCS.*`.\n")
    print("| kernel | " + " | ".join(KINDS) + " |")
    print("|---|" + "---|" * len(KINDS))
    total = collections.Counter()
    for fn, c in counts.items():
        if not c:
            continue
        total.update(c)
        print(f"| `{fn[:110]}` | " + " | ".join(str(c.get(k, "")) for k in KINDS) + " |")
    print("| **library total** | " + " | ".join(str(total.get(k, 0)) for k in KINDS) + " |\n")
    print("First instruction of each kind in the A-stationary fused GEMM and the CTA-pair attention kernel:\n\n```")
    for (fn, k), body in samples.items():
        print(f"{fn.split('::')[-1][:48]:<48} {body}")
    print("```")


if __name__ == "__main__":
    main()
