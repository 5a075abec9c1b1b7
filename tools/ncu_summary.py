"""Text summary of an .ncu-rep (run here, no GPU needed): key raw metrics per kernel + hot SASS regions."""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'dram__bytes_read.sum [',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active',
        'sm__inst_executed_pipe_tensor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread ', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ', 'lts__t_bytes.sum ', 'smsp__inst_executed.sum ',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max ']


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(path):
    raw = list(csv.reader(io.StringIO(run(["-i", path, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        name = r[hdr.index('Kernel Name')]
        print(f"=== {name}")
        for i, h in enumerate(hdr):
            hh = h + ' '
            if any(k in hh or k in h + ' [' for k in KEYS):
                if 'stalled' in h:
                    try:
                        if float(r[i]) < 0.2:
                            continue
                    except ValueError:
                        continue
                print(f"  {h} [{units[i]}] = {r[i]}")
    src = list(csv.reader(io.StringIO(run(["-i", path, "--page", "source", "--csv"]))))
    # the source page concatenates kernels: split on header rows
    blocks, cur = [], None
    for row in src:
        if row and row[0] == 'Kernel Name':
            cur = dict(name=row[1], rows=[])
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(row)
    for b in blocks:
        rows = b['rows']
        if len(rows) < 3:
            continue
        hdr = rows[0]
        data = [r for r in rows[1:] if len(r) == len(hdr)]
        try:
            ie, sm, so = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
        except ValueError:
            continue
        tot = sum(float(r[ie]) for r in data) or 1.0
        ts = sum(float(r[sm]) for r in data) or 1.0
        print(f"=== hot SASS regions of {b['name']}: {len(data)} SASS instructions, {tot / 1e9:.3f} G warp-instructions")
        runs, c = [], None
        for i, r in enumerate(data):
            e = float(r[ie])
            if c and abs(e - c['e']) <= 0.03 * max(e, c['e'], 1):
                c['n'] += 1; c['sum'] += e; c['samples'] += float(r[sm]); c['end'] = i
            else:
                if c:
                    runs.append(c)
                c = dict(start=i, end=i, e=e, n=1, sum=e, samples=float(r[sm]))
        runs.append(c)
        for c in sorted([c for c in runs if c['sum'] / tot > 0.02 or c['samples'] / ts > 0.03], key=lambda c: c['start']):
            print(f"  sass[{c['start']:5d}..{c['end']:5d}] n={c['n']:4d} exec/instr={c['e'] / 1e6:9.2f}M "
                  f"instr share={100 * c['sum'] / tot:5.1f}% stall-sample share={100 * c['samples'] / ts:5.1f}%  "
                  f"{data[c['start']][so].strip()[:48]}")


if __name__ == "__main__":
    main(sys.argv[1])
