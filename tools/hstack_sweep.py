import sys, json, itertools
sys.path.insert(0,'/root/repo/tools'); sys.path.insert(0,'/root/repo')
from conv_sweep import run, LAYERS
layers = sys.argv[1].split(',')
hss = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 1]
for name in layers:
    for hs, eg, ab in itertools.product([0],[2,4],[2]):
        ov=dict(hstack=hs, epi_groups=eg, acc_bufs=ab)
        try:
            r,by,fl = run(name, 32, ov, 10, 'fp16')
        except Exception as e:
            print(name, ov, 'ERR', str(e)[:80]); continue
        p=r['plan']; ms=r['ms']
        print(f"{name} hs={hs} G={eg} bufs={ab} ms={ms:.4f} {by/ms/1e6:.0f} GB/s | TH={p['TH']} mt={p['n_mtiles']} Nt={p['N_tile']} st={p['stages']} nk={p['n_k']} tmem={p['tmem_cols']} smem={p['smem_bytes']}", flush=True)
