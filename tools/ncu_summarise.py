#!/usr/bin/env python
"""Developer tool: turns the ncu captures of one gpurun call into the small tracked files under profiles/.

  python tools/ncu_summarise.py launches gpurun_out/cNN/launches.csv profiles/rRR_final      # -> _launches.csv, _launch_shares.csv, gemm traffic json
  python tools/ncu_summarise.py full     gpurun_out/cNN/top.ncu-rep  profiles/rRR_final_ncu_full_summary.csv
The launch list is `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` of `bench.py --steps 1
--warmup 1`: the last COMPLETE step of the capture is kept (a step = the launches from one mel kernel to the next)."""
import csv, json, re, subprocess, sys

KEEP = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__cluster_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']


def short(name):
    name = re.sub(r'^void ', '', name).replace('some::', '')
    return re.sub(r'\(.*', '', name)


def launches(src, prefix):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith('=='))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        e = per.setdefault(int(r[ix['ID']]), {'kernel': r[ix['Kernel Name']]})
        val = float(r[ix['Metric Value']].replace(',', ''))
        unit = r[ix['Metric Unit']]
        m = r[ix['Metric Name']]
        if m == 'gpu__time_duration.sum':
            e['us'] = val * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3, 'second': 1e6}.get(unit, 1.0)
        else:
            scale = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(unit, 1e-6)
            e['rd' if 'read' in m else 'wr'] = val * scale
    ids = sorted(per)
    ours = [i for i in ids if re.search(r'some::|gemm_|attention_tc|layernorm|dwconv|mel_kernel|decode_|bound_head', per[i]['kernel'])]
    mel = [n for n, i in enumerate(ours) if 'mel_kernel' in per[i]['kernel']]
    segs = [ours[a:b] for a, b in zip(mel, mel[1:] + [len(ours)])]      # one segment per step: mel .. decode
    full_len = max(len(s_) for s_ in segs)
    step = [s_ for s_ in segs if len(s_) == full_len][-1]               # -c N may cut the capture inside the last step
    with open(prefix + '_launches.csv', 'w') as f:
        f.write('index,kernel,us,dram_read_MB,dram_write_MB\n')
        for n, i in enumerate(step):
            e = per[i]
            f.write(f'{n},"{e["kernel"][:90]}",{e["us"]:.2f},{e.get("rd", 0):.1f},{e.get("wr", 0):.1f}\n')
    agg = {}
    for i in step:
        e = per[i]
        k = short(e['kernel'])
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += e['us']
        a[2] += (e.get('rd', 0) + e.get('wr', 0)) / 1e3
    tot = sum(a[1] for a in agg.values())
    with open(prefix + '_launch_shares.csv', 'w') as f:
        f.write('kernel,launches,us,share,dram_GB\n')
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{k},{a[0]},{a[1]:.1f},{a[1] / tot:.4f},{a[2]:.3f}\n')
        f.write(f'TOTAL,{len(step)},{tot:.1f},1.0000,{sum(a[2] for a in agg.values()):.3f}\n')
    g = [per[i] for i in step if 'gemm' in per[i]['kernel']]
    byts = sum((e.get('rd', 0) + e.get('wr', 0)) * 1e6 for e in g)
    print(f'{len(step)} launches in the step, {tot / 1e3:.2f} ms of kernels (cold-cache, serialised); gemm: {len(g)} launches, {byts / 1e9:.2f} GB DRAM')
    return {'gemm_launches': len(g), 'dram_bytes_per_step': byts, 'dram_bytes_per_launch': byts / max(1, len(g))}


def full(rep, dst):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(dst, 'w') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + [f'launch{i}' for i in range(len(data))])
        w.writerow(['Kernel Name', ''] + [re.sub(r'^void ', '', r[ix['Kernel Name']]).replace('some::', '')[:70] for r in data])
        for m in KEEP:
            if m in ix:
                w.writerow([m, units[ix[m]]] + [r[ix[m]] for r in data])
    print(f'{len(data)} launches -> {dst}')


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        info = launches(sys.argv[2], sys.argv[3])
        print(json.dumps(info))
    else:
        full(sys.argv[2], sys.argv[3])
