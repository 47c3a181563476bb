"""In-order issue model of one loop body of a kernel's SASS: which instructions does a lone warp wait on?
usage: python tools/sass_inorder_sim.py <lib.so> <kernel substring> <start addr hex> <end addr hex> [iterations]
Latencies measured with tools/microbench_lat.cu on B200."""
import re, subprocess, sys
so, kern, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 4
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
on = False; ins = []
for line in txt.splitlines():
    if 'Function :' in line: on = kern in line
    if not on: continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
    if m:
        a = int(m.group(1), 16)
        if lo <= a <= hi: ins.append((a, m.group(2).strip()))
LAT = {'FADD': 4, 'FMUL': 4, 'FFMA': 4, 'HFMA2': 4, 'IMAD': 4, 'SHFL': 26, 'LDS': 29, 'MUFU': 18, 'LDG': 300}
def lat(op):
    for k, v in LAT.items():
        if op.startswith(k): return v
    return 6   # alu pipe / cross-pipe
def parse(text):
    pred = None
    m = re.match(r'@(!?U?P\d)\s+(.*)', text)
    if m: pred, text = m.group(1).lstrip('!'), m.group(2)
    op, _, rest = text.partition(' ')
    regs = re.findall(r'\b(U?R\d+|U?P\d)\b', rest)
    ops = [o.strip() for o in rest.split(',')]
    dst = []
    if op.startswith(('ST', 'BAR', 'BRA', 'BSSY', 'BSYNC', 'NOP', 'WARPSYNC', 'EXIT')): dst = []
    else:
        d0 = re.findall(r'\b(U?R\d+|U?P\d)\b', ops[0]) if ops else []
        dst = d0[:1]
        if op.startswith(('ISETP', 'FSETP')) and len(ops) > 1:
            dst += re.findall(r'\b(P\d)\b', ops[1])
        if '.64' in op and dst and dst[0].startswith('R'): dst.append('R%d' % (int(dst[0][1:]) + 1))
    src = [r for r in regs if r not in dst[:1]] + ([pred] if pred else [])
    if op.startswith(('ST',)): src = regs
    return op, dst, src
ready = {}; T = 0; stall_by = {}; first = None
for it in range(iters):
    t_start = T
    for a, text in ins:
        if any(s in text for s in ('__frcp', 'CALL')): continue
        op, dst, src = parse(text)
        if op.startswith(('BRA', 'BSSY', 'BSYNC', 'WARPSYNC', 'NOP')) and 'DIV' not in op:
            T += 1; continue
        t_ready = max([ready.get(r, 0) for r in src] + [T])
        if it == iters - 1 and t_ready > T:
            stall_by[(a, text)] = t_ready - T
        T = t_ready + 1
        L = lat(op)
        for d in dst: ready[d] = T - 1 + L
        if op.startswith('BAR'): T += 16
    if it == iters - 1:
        print(f"{len(ins)} instructions, {T - t_start} cycles for the last pass")
for (a, text), s in sorted(stall_by.items(), key=lambda kv: -kv[1])[:40]:
    print(f"  {a:05x} waits {s:3d}  {text[:80]}")
