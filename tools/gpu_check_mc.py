import sys, os, time, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
from common import *
import skirt_b200 as sk
B = 8
cfg = cfg_c1(n=40, packages=2e5, threads=8, storeabs=1)
S = make_ref(cfg).setup()
Npp = S.packages_per_lambda(); print('ref Npp', Npp)
t = S.grid_tables(); med = S.medium()
e = setup_engine(sk.Engine(0), cfg, t, med, S.luminosities())
ref_f, ref_s, ref_l, gpu_f, gpu_s, gpu_l = [],[],[],[],[],[]
tref=0; tgpu=0
for b in range(B):
    S.reset(1000+b*17); tref += S.run_stellar(); ins = S.instruments()
    ref_f.append(ins[0]['frame'].copy()); ref_s.append(ins[1]['sed'].copy()); ref_l.append(S.labs().copy())
    e.reset_results(); st = e.run_stellar(Npp, store_absorption=True, seed=77+b); tgpu += st['kernel_ms']/1e3
    gpu_f.append(e.fetch_frame(0)); gpu_s.append(e.fetch_sed(1)); gpu_l.append(e.fetch_labs())
print('stats', st)
print('ref %.3g pkt/s (8 thr)  gpu %.3g pkt/s' % (B*Npp/tref, B*Npp/tgpu))
def cmp(name, a, b):
    a=np.array(a); b=np.array(b)
    ma, mb = a.mean(0), b.mean(0); va, vb = a.var(0, ddof=1)/len(a), b.var(0, ddof=1)/len(b)
    sig = np.sqrt(va+vb); ok = sig>0
    z = (ma-mb)[ok]/sig[ok]
    tot_a, tot_b = a.reshape(len(a),-1).sum(1), b.reshape(len(b),-1).sum(1)
    zt = (tot_a.mean()-tot_b.mean())/np.sqrt(tot_a.var(ddof=1)/len(a)+tot_b.var(ddof=1)/len(b))
    print('%-6s total ref %.6g gpu %.6g  z_total %.2f | bins %d  |z|<3: %.4f  mean z %.3f  rms z %.3f' % (name, tot_a.mean(), tot_b.mean(), zt, ok.sum(), (np.abs(z)<3).mean(), z.mean(), z.std()))
cmp('frame', ref_f, gpu_f); cmp('sed', ref_s, gpu_s); cmp('labs', ref_l, gpu_l)
