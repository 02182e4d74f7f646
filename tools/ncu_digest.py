#!/usr/bin/env python
"""Digest of an .ncu-rep: headline metrics + hottest source lines (stall samples, instructions)."""
import csv, subprocess, sys, io

KEYS = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum ', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread ', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg ',
        'smsp__average_warps_issue_stalled', 'launch__shared_mem_per_block_dynamic']

def run(args):
    return subprocess.run(['ncu', '-i'] + args, capture_output=True, text=True).stdout

def main(path, top=30):
    raw = list(csv.reader(io.StringIO(run([path, '--page', 'raw', '--csv']))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(k.strip()) if k.endswith(' ') else k in h for k in KEYS):
            if 'stalled' in h and float(v.replace(',', '') or 0) < 0.3:
                continue
            print(f"{h:95s} {v} {u}")
    rows = list(csv.reader(io.StringIO(run([path, '--page', 'source', '--csv', '--print-source', 'cuda,sass']))))
    cur, out, tot, toti = None, [], 0, 0
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]; continue
        if len(r) < 8 or r[0] in ('Line No', 'Function Name', ''):
            continue
        try:
            ln, s, ie = int(r[0]), int(r[4]), int(r[7])
        except ValueError:
            continue
        out.append((s, ie, cur, ln, r[1].strip()[:100])); tot += s; toti += ie
    out.sort(reverse=True)
    print(f"total samples {tot}  warp-instructions {toti/1e6:.1f}M")
    for s, ie, f, ln, src in out[:top]:
        print(f"{100*s/max(tot,1):5.1f}% inst={ie/1e6:7.2f}M {f}:{ln}: {src}")

def to_json(path, key, digest_txt, workload, out='profiles/ncu_digest.json'):
    """Record the numbers bench.py quotes (DRAM bytes per launch, tensor-pipe activity) for kernel `key` in profiles/ncu_digest.json."""
    import json, os
    raw = list(csv.reader(io.StringIO(run([path, '--page', 'raw', '--csv']))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    get = lambda name: next((float(v.replace(',', '')) * {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1}.get(u, 1)
                             for h, u, v in zip(hdr, units, vals) if h == name), None)
    d = json.load(open(out)) if os.path.isfile(out) else {}
    d[key] = {'dram_bytes': (get('dram__bytes_read.sum') or 0) + (get('dram__bytes_write.sum') or 0),
              'tensor_pipe_active_pct': get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
              'xu_pipe_pct': get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
              'duration_us': next((float(v.replace(',', '')) * {'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6}.get(u, 1.0)
                                  for h, u, v in zip(hdr, units, vals) if h == 'gpu__time_duration.sum'), None), 'workload': workload, 'source': digest_txt}
    json.dump(d, open(out, 'w'), indent=1)


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[2] == '--json':     # ncu_digest.py rep --json key digest.txt "workload"
        to_json(sys.argv[1], sys.argv[3], sys.argv[4], sys.argv[5], *(sys.argv[6:7]))
    else:
        main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
