"""Summarise an .ncu-rep: key raw metrics per kernel + SASS hot-spot segments (run on the CPU box)."""
import csv, io, subprocess, sys

rep = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'sm__cycles_active.avg', 'lts__t_bytes.sum']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    print('====', name[:110])
    for w in WANT:
        if w in hdr:
            print(f'  {w:70s} {r[hdr.index(w)]:>18s} {rows[1][hdr.index(w)]}')


def hot(rep, kernel_regex, top=18):
    import re
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name"')
    for bi, blk in enumerate(blocks[1:]):
        if bi % 2 == 1:
            continue  # every kernel is listed twice
        rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
        if not re.search(kernel_regex, rows[0][1]):
            continue
        print('==== HOT', rows[0][1][:100])
        hdr = rows[1]
        isrc = hdr.index('Source'); ist = hdr.index('Warp Stall Sampling (All Samples)')
        iex = hdr.index('Instructions Executed'); ith = hdr.index('Avg. Threads Executed')
        seen = set(); data = []
        for r in rows[2:]:
            if len(r) <= ist or r[0] in seen:
                continue
            seen.add(r[0])
            try:
                data.append((int(r[ist] or 0), int(r[iex] or 0), r[isrc][:95], r[ith]))
            except ValueError:
                pass
        tot = sum(d[0] for d in data) or 1
        toti = sum(d[1] for d in data) or 1
        print(f'  samples {tot} instr {toti} sass {len(data)}')
        for i, d in sorted(enumerate(data), key=lambda x: -x[1][0])[:top]:
            print(f'  [{i:4d}] {100*d[0]/tot:5.1f}% smp  {100*d[1]/toti:5.1f}% ins  thr={d[3]:>4s}  {d[2]}')


if len(sys.argv) > 2:
    hot(rep, sys.argv[2])
