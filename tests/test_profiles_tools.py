"""The measurement tooling stays runnable: the committed ncu launch list parses into the per-kernel tables that DESIGN.md and
bench.py quote (tools/launch_table.py, tools/dram_table.py), and bench.py finds the DRAM-traffic table."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_table_matches_committed_summary():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'launch_table.py'), os.path.join(ROOT, 'profiles', 'r02_launches.csv')],
                         capture_output=True, text=True, check=True).stdout
    rows = {l.split()[0]: l.split() for l in out.splitlines() if len(l.split()) == 4 and l.split()[1].isdigit()}
    assert int(rows['conv_gemm_kernel'][1]) == 99                      # conv launches of one forward
    assert sum(int(r[1]) for r in rows.values()) == 198                 # all launches of one forward (185 + 13 split-K reductions)
    committed = open(os.path.join(ROOT, 'profiles', 'r02_launches_summary.txt')).read()
    assert rows['conv_gemm_kernel'][2] in committed                     # same total time as the committed summary


def test_dram_table_and_bench_lookup(tmp_path):
    dst = tmp_path / 'dram.json'
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'dram_table.py'), os.path.join(ROOT, 'profiles', 'r02_launches.csv'), str(dst)],
                   capture_output=True, text=True, check=True)
    got = json.load(open(dst))['kernels']
    ref = json.load(open(os.path.join(ROOT, 'profiles', 'r02_dram_traffic.json')))['kernels']
    assert got['conv_gemm_kernel']['launches'] == ref['conv_gemm_kernel']['launches'] == 99
    assert abs(got['conv_gemm_kernel']['dram_bytes_per_launch'] - ref['conv_gemm_kernel']['dram_bytes_per_launch']) < 1.0
    sys.path.insert(0, ROOT)
    import bench
    tr, src = bench._traffic('conv_gemm_kernel')
    assert tr == ref['conv_gemm_kernel']['dram_bytes_per_launch'] and 'r02_dram_traffic.json' in src
    assert bench._traffic('no_such_kernel') == (None, None)


def test_every_launched_kernel_has_a_committed_ncu_summary():
    """north_star: every kernel evidenced by a committed ncu capture.  Every kernel of the launch list of one forward appears in an
    `ncu --set full` summary under profiles/; the one kernel added after the last full capture is evidenced by the launch list itself
    (time + DRAM bytes per launch) and must then appear in the DRAM table."""
    import glob
    names = [l.split()[0] for l in open(os.path.join(ROOT, 'profiles', 'r02_launches_summary.txt')) if len(l.split()) == 4 and l.split()[1].isdigit()]
    assert len(names) >= 20
    full = ''.join(open(f).read() for f in glob.glob(os.path.join(ROOT, 'profiles', 'r02_ncu_*.txt')))
    dram = json.load(open(os.path.join(ROOT, 'profiles', 'r02_dram_traffic.json')))['kernels']
    launch_list_only = {'fir_down_stream_kernel'}
    for n in names:
        base = n.split('<')[0]
        if base in launch_list_only:
            assert base in dram and dram[base]['dram_bytes_per_launch'] > 0
        else:
            assert base in full, f'no ncu --set full summary for {n}'
