"""Summarises an ncu report: key raw metrics + per-source-line hotspots of one kernel.
usage: python scripts/ncu_summary.py <prof.ncu-rep> <cubin-or-.so> <mangled-kernel-prefix> <out-prefix>"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, binary, kprefix, outp = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
pats = ['gpu__time_duration', 'launch__', 'sm__throughput', 'smsp__issue_active', 'sm__inst_executed_pipe_', 'sm__pipe_',
        'smsp__thread_inst_executed_per_inst', 'dram__bytes', 'lts__t_bytes.sum', 'smsp__inst_executed.sum',
        'smsp__sass_thread_inst_executed_op_', 'sm__warps_active', 'smsp__warps_eligible', 'sm__cycles_elapsed.avg',
        'smsp__cycles_active.avg', 'warp_issue_stalled', 'l1tex__t_bytes.sum', 'smsp__sass_average_branch',
        'l1tex__data_bank_conflicts', 'smsp__average_warp']
with open(outp + "_metrics.csv", "w") as f:
    w = csv.writer(f); w.writerow(["metric", "unit", "value"])
    for k in hdr:
        if any(p in k for p in pats): w.writerow([k, u[k], d[k]])
for k in ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'smsp__warps_eligible.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
          'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
          'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum']:
    if k in d: print(f"{k:80s} {d[k]:>18s} {u[k]}")

td = tempfile.mkdtemp()
if binary.endswith(".so"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binary)], cwd=td, capture_output=True)
    cubins = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")]
else:
    cubins = [binary]
sass = None
for cb in cubins:
    out = subprocess.run(["nvdisasm", "-gi", cb], capture_output=True, text=True).stdout.splitlines()
    st = [i for i, l in enumerate(out) if l.startswith(".text." + kprefix)]
    if st:
        en = [i for i, l in enumerate(out) if i > st[0] and l.strip().startswith(".section")]
        sass = out[st[0]:(en[0] if en else len(out))]
        break
# nvdisasm -gi prints, before an instruction, the inline chain leaf first: 'File F, line N inlined at G, line M'
# then the frame (G, M) itself ... down to the line of the kernel body; chain[0] = leaf, chain[-1] = root
cur, chain, off2, offchain, fresh = None, [], {}, {}, True
for ln in sass:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh: chain = []; fresh = False
        chain.append((m.group(1).split('/')[-1], int(m.group(2))))
        cur = chain[0]
        continue
    m = re.match(r'\s*(\$[\w$.]+):\s*$', ln)
    if m:  # a subroutine embedded in the kernel (IEEE slow paths, noinline device functions): most carry no line info
        name = m.group(1)
        short = re.sub(r'^\$__internal_\d+_\$', '', name) if name.startswith('$__internal') else name.split('$')[-1][:60]
        chain = [('<' + short + '>', 0)]
        cur = chain[0]
        fresh = True
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
    if m:
        off2[int(m.group(1), 16)] = (cur, m.group(2).strip()); offchain[int(m.group(1), 16)] = tuple(chain); fresh = True
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]; ia, ii, it, isamp = h.index('Address'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
base = None; by = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]
bysite = [collections.defaultdict(lambda: [0, 0, 0]) for _ in range(3)]  # frames counted from the kernel body
for r in rows[2:]:
    a = int(r[ia], 16)
    if base is None: base = a
    cur, txt = off2.get(a - base, (None, r[1]))
    key = cur if cur else ('?', 0)
    inst, th, sm = int(r[ii]), int(r[it]), int(r[isamp])
    b = by[key]; b[0] += inst; b[1] += th; b[2] += sm
    ch = [c for c in offchain.get(a - base, ()) if c[0].endswith('.cuh') or c[0].endswith('.cu') or c[0].startswith('<')]
    for lvl in range(3):
        k2 = tuple(reversed(ch[-(lvl + 1):])) if ch else (('?', 0),)
        b = bysite[lvl][k2]; b[0] += inst; b[1] += th; b[2] += sm
    tot[0] += inst; tot[1] += th; tot[2] += sm
cache = {}
def srcline(f, l):
    if f not in cache:
        for base_dir in ("cpuperformanceraytracer_b200/csrc", "."):
            p = os.path.join(base_dir, f)
            if os.path.exists(p): cache[f] = open(p).read().splitlines(); break
        else: cache[f] = []
    s = cache[f]; return s[l - 1].strip()[:88] if 0 < l <= len(s) else ''
with open(outp + "_hotspots.txt", "w") as f:
    f.write('total warp-level instructions %.4e, thread-level %.4e, avg active threads/inst %.2f, samples %d\n' % (tot[0], tot[1], tot[1] / tot[0], tot[2]))
    f.write('by source line (share of warp-level instructions issued, avg active threads, share of stall samples)\n')
    for key, b in sorted(by.items(), key=lambda kv: -kv[1][0])[:60]:
        f.write('%-16s:%4d inst%%=%5.2f eff=%5.1f samp%%=%5.2f | %s\n' % (key[0], key[1], 100 * b[0] / tot[0], b[1] / max(b[0], 1), 100 * b[2] / max(tot[2], 1), srcline(*key)))
    for lvl in range(3):
        f.write('\nby call site, %d frame(s) below the kernel body (root first)\n' % lvl)
        for key, b in sorted(bysite[lvl].items(), key=lambda kv: -kv[1][0])[:(20, 45, 70)[lvl]]:
            f.write('%-34s inst%%=%5.2f eff=%5.1f samp%%=%5.2f | %s\n' % ('>'.join(('%d' % k[1]) if not k[0].startswith('<') else k[0] for k in key), 100 * b[0] / tot[0], b[1] / max(b[0], 1),
                                                                       100 * b[2] / max(tot[2], 1), srcline(*key[-1])))
print(open(outp + "_hotspots.txt").read())
