"""Summarise an .ncu-rep: per kernel key metrics + top stalled SASS instructions (reads with `ncu -i`)."""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 22
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp16.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum',
        'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__grid_size', 'launch__block_size',
        'smsp__cycles_active.avg', 'sm__cycles_active.avg']
seen = set()
for r in rows[2:]:
    name = r[idx['Kernel Name']]
    if name in seen:
        continue
    seen.add(name)
    print('=====', name[:100])
    for w in want:
        if w in idx:
            print(f"  {w:72s} {r[idx[w]]} {units[idx[w]]}")
if kern:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data, seen_addr = [], set()
    for r in rows[2:]:
        if len(r) < len(hdr) - 2 or not r[0].startswith('0x'):
            continue
        if r[0] in seen_addr:
            break
        seen_addr.add(r[0])
        data.append(r)
    I = lambda r, h: int(r[idx[h]] or 0)
    tot = sum(I(r, '# Samples') for r in data)
    print('SASS instrs', len(data), 'samples', tot)
    op, samp = collections.Counter(), collections.Counter()
    for r in data:
        s = r[idx['Source']].strip()
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', s)
        o = m.group(2).split('.')[0] if m else s
        op[o] += I(r, 'Instructions Executed')
        samp[o] += I(r, '# Samples')
    print('executed by opcode:', [(o, c, samp[o]) for o, c in op.most_common(16)])
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    ss = {h: sum(I(r, h) for r in data) for h in stalls}
    print('stalls:', sorted(ss.items(), key=lambda x: -x[1])[:8])
    for r in sorted(data, key=lambda r: -I(r, '# Samples'))[:topn]:
        st = {h: I(r, h) for h in stalls if I(r, h) > 0}
        print(f"  {r[idx['# Samples']]:>5} {r[idx['Instructions Executed']]:>9} {r[idx['Source']].strip()[:78]:78s}",
              sorted(st.items(), key=lambda x: -x[1])[:2])
