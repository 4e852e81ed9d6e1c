"""profiles/r2_ncu_raw.csv (ncu --set full raw page of tools/prof_round2.py) -> profiles/r2_traffic.json:
DRAM bytes per launch, ncu duration, tensor-pipe and issue activity of the hot kernels (read by
bench.py for `roofline.traffic`).   python tools/make_traffic_json.py <raw.csv> <out.json>"""
import csv
import json
import sys


def num(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return None


def main(raw, out):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}

    def bytes_of(r, name):
        v = num(r[col[name]])
        return None if v is None else v * scale[units[col[name]]]

    pick = {'sigma': 'sigma_tc_kernel', 'lvis': 'lvis_tc3_kernel<0, 0>', 'point': 'point_tc_kernel',
            'integrate': 'integrate_kernel<1, 0>', 'lvis_front_lit_chunk': 'lvis_tc3_kernel<0, 1>'}
    kernels = {}
    for key, pat in pick.items():
        for r in data:
            name = r[col['Kernel Name']]
            rd, wr = bytes_of(r, 'dram__bytes_read.sum'), bytes_of(r, 'dram__bytes_write.sum')
            if pat in name and rd is not None and rd == rd and key not in kernels:
                if key == 'lvis' and wr < 1e8:          # the one-kernel fused variant writes no lvis
                    continue
                kernels[key] = {
                    'kernel': name.replace('void <unnamed>::', '').replace('<unnamed>::', '').split('(')[0],
                    'dram_bytes': int(rd + wr), 'dram_read': int(rd), 'dram_write': int(wr),
                    'ncu_ms': num(r[col['gpu__time_duration.sum']]),
                    'tensor_pipe_active_pct': num(r[col['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]),
                    'issue_active_pct': num(r[col['smsp__issue_active.avg.pct_of_peak_sustained_active']]),
                    'sm_clock_ghz': num(r[col['sm__cycles_elapsed.avg.per_second']])}
    json.dump({'workload': {'imh': 800, 'imw': 800, 'spp': 128, 'light_dirs': 512},
               'source': 'profiles/r2_ncu_full.csv (ncu --set full --clock-control none of '
                         'tools/prof_round2.py, one launch each)',
               'kernels': kernels}, open(out, 'w'), indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
