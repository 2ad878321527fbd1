"""Opcode evidence for the tensor-core matchers, from the SASS of the built objects (run here,
no GPU needed):

    python scripts/sass_tc_summary.py [tag]      ->  profiles/<tag>_hamming_tc_sass.md

Per kernel of hamming_tc4.o / hamming_tc.o: the counts of the Blackwell-only mnemonics
(B200_PROFILING.md's list: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit,
UTCATOMSWS = tcgen05.alloc/dealloc, UBLKCP = cp.async.bulk, SYNCS = mbarrier) and the full histogram
of the headline kernel.  The per-item arithmetic the roofline line in bench.py relies on
(4 UTCOMMA of K = 64 per 128 x 128 item for format 4, 8 UTCIMMA of K = 32 for format 8) can be
read off the UTC*MMA lines listed at the end."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJDIR = os.path.join(ROOT, "pybot_b200", "csrc", "_obj")
CSRC = os.path.join(ROOT, "pybot_b200", "csrc")
OBJS = [("hamming_tc4", "hamming_tc4_kernel"), ("hamming_tc", "hamming_tc_kernel")]
MARK = ["UTCOMMA", "UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG",
        "UBLKCP", "UBLKPF", "SYNCS", "ELECT", "REDG", "FMNMX3", "VIMNMX3", "VIMNMX", "POPC", "HMMA", "IMMA"]


def kernels(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, name = collections.OrderedDict(), None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            out[name] = []
            continue
        m = re.match(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m and name:
            out[name].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(txt):
    return re.sub(r"^@!?U?P\w+\s+", "", txt).split()[0]


def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True)
    return r.stdout.strip().split("(")[0] if r.returncode == 0 else n


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    lines = ["# %s: SASS evidence for the tensor-core matchers" % tag, "",
             "Written by `scripts/sass_tc_summary.py` from `cuobjdump -sass` of the objects that are linked into",
             "`libpybot_b200.so` (sm_100a only).  Sources: " + ", ".join(
                 "`%s.cu` sha256 %s" % (b, hashlib.sha256(open(os.path.join(CSRC, b + ".cu"), "rb").read()).hexdigest()[:16])
                 for b, _ in OBJS) + ".", "",
             "| kernel | instructions | " + " | ".join(MARK) + " |", "|---|---:|" + "---:|" * len(MARK)]
    detail = []
    for base, head in OBJS:
        ks = kernels(os.path.join(OBJDIR, base + ".o"))
        for name, ins in ks.items():
            full = collections.Counter(opcode(t) for _, t in ins)
            fam = collections.Counter()
            for op, c in full.items():
                fam[op.split(".")[0]] += c
            lines.append("| `%s` | %d | %s |" % (demangle(name), len(ins), " | ".join(str(fam.get(m, 0) or "") for m in MARK)))
            if head in name and "peak" not in name and "expand" not in name:
                detail.append((demangle(name), ins, full, fam))
    lines += ["", "No `HMMA` / `IMMA` (legacy `mma.sync`) and no `POPC` in either tensor-core kernel: every descriptor pair is",
              "scored by `tcgen05.mma` (`UTC*MMA`), accumulators are read from TMEM with `tcgen05.ld` (`LDTM`), the operand",
              "ring is filled with `cp.async.bulk` (`UBLKCP`) and handed over through mbarriers (`SYNCS`).", ""]
    for name, ins, full, fam in detail:
        lines += ["## `%s`" % name, "", "Full-mnemonic histogram (top 40 of %d instructions):" % len(ins), "",
                  "| opcode | count |", "|---|---:|"]
        for op, c in full.most_common(40):
            lines.append("| `%s` | %d |" % (op, c))
        lines += ["", "Tensor-core, TMEM and bulk-copy instructions in program order:", "", "```"]
        for a, t in ins:
            if re.search(r"UTC\w*MMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UBLKCP|UTMALDG", t):
                lines.append("/*%04x*/  %s ;" % (a, t))
        lines += ["```", ""]
    out = os.path.join(ROOT, "profiles", "%s_hamming_tc_sass.md" % tag)
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(out)


if __name__ == "__main__":
    main()
