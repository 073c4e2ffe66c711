"""Opcode histogram of the shipped library's SASS, per kernel: the evidence that the hot path is tcgen05 / TMEM / TMA code.

    python tools/sass_opcodes.py [infur_b200/lib/libinfur_b200.so] > profiles/rN_sass_opcodes.txt

Counts, per kernel (demangled name, template arguments kept), the mnemonics B200_PROFILING.md names: UTCHMMA / UTCIMMA
(tcgen05.mma kind::f16 / kind::i8, `.2CTA` = cta_group::2), UTMALDG / UTMASTG (TMA load / store), LDTM (tcgen05.ld), UTCBAR
(tcgen05.commit), SYNCS (mbarrier), and the legacy tensor-core opcodes HMMA / IMMA that must NOT appear.
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "infur_b200/lib/libinfur_b200.so"
WATCH = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "IMMA", "FADD2", "FMUL2", "FHADD",
         "LDG", "STG", "LDS", "STS", "CCTL", "MEMBAR", "ERRBAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
    kernels = collections.OrderedDict()
    cur = None
    it = iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(it, m.group(1))
            cur = re.sub(r"^void ", "", cur).replace("infur::<unnamed>::", "").replace("infur::(anonymous namespace)::", "")
            if "<" in cur:   # keep the template arguments, drop the parameter list
                depth, end = 0, len(cur)
                for i, ch in enumerate(cur):
                    depth += ch == "<"
                    depth -= ch == ">"
                    if ch == ">" and depth == 0:
                        end = i + 1
                        break
                cur = cur[:end]
            else:
                cur = cur.split("(")[0]
            cur = cur.replace("(int)", "").replace("(bool)", "")
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op, mods = m.group(1), m.group(2)
            kernels[cur]["__total__"] += 1
            if op in WATCH:
                kernels[cur][op] += 1
                if ".2CTA" in mods:
                    kernels[cur][op + ".2CTA"] += 1
                if op == "UTCBAR" and "MULTICAST" in mods:
                    kernels[cur]["UTCBAR.MULTICAST"] += 1
    print(f"SASS opcode histogram of {LIB} (cuobjdump -sass; sm_100a only)\n")
    tot = collections.Counter()
    for k, c in kernels.items():
        tot.update(c)
        keys = [x for x in c if x != "__total__"]
        body = "  ".join(f"{x} {c[x]}" for x in sorted(keys))
        print(f"{k}\n    instructions {c['__total__']}  |  {body}\n")
    print("TOTAL  " + "  ".join(f"{x} {tot[x]}" for x in sorted(tot) if x != "__total__"))
    legacy = tot["HMMA"] + tot["IMMA"]
    print(f"\nlegacy tensor-core opcodes (HMMA / IMMA, i.e. mma.sync / wmma): {legacy}")
    archs = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    print("embedded cubins: " + ", ".join(sorted(set(re.findall(r"sm_\d+a?", archs)))))


if __name__ == "__main__":
    main()
