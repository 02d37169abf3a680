"""Per-kernel SASS mnemonic counts of liblm_bev.so (the evidence file profiles/r02_sass_kernels.txt):
   python tools/sass_summary.py [path/to/liblm_bev.so] > profiles/r02_sass_kernels.txt"""
import collections, os, re, subprocess, sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "lanemapping_b200", "csrc", "liblm_bev.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
SPECIAL = re.compile(r"^(UBLKCP|UBLKPF|SYNCS|FFMA2|FMUL2|FADD2|ATOMS|ATOMG|ATOM|RED|REDG|REDUX|VOTEU|LDGSTS|LDGDEPBAR|PRMT)")
MEM = re.compile(r"^(LDG|STG|LDS|STS|LDL|STL|BAR)")
print("""SASS evidence per kernel of liblm_bev.so (cuobjdump -sass, sm_100a): instruction counts by mnemonic.
UBLKCP = TMA 1-D bulk copy (cp.async.bulk), UBLKPF = L2 bulk prefetch (cp.async.bulk.prefetch.L2), SYNCS.* = mbarrier
arrive.expect_tx / try_wait, FFMA2/FMUL2/FADD2 = packed FP32x2 (Blackwell), ATOMS/ATOMG/RED* = shared/global atomics,
REDUX = warp reductions, LDGSTS = cp.async (global -> shared, no register), PRMT = byte permute.
Regenerate: python tools/sass_summary.py > profiles/r02_sass_kernels.txt
""")
for block in sass.split("Function : ")[1:]:
    name = block.split("\n", 1)[0].strip()
    try:
        name = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    except FileNotFoundError:
        pass
    name = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    m = re.match(r"^(\w+(?:<[^>]*>)?)", name)                   # kernel name + template arguments, no parameter list
    name = m.group(1) if m else name
    ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", block)
    special, mem = collections.Counter(), collections.Counter()
    for op in ops:
        if SPECIAL.match(op):
            special[op] += 1
        elif MEM.match(op):
            mem[op] += 1
    fmt = lambda c: ", ".join(f"{k} x{v}" for k, v in sorted(c.items())) or "-"
    print(f"== {name}  ({len(ops)} SASS instructions)\n   {fmt(special)}\n   mem: {fmt(mem)}")
