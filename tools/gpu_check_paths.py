import sys, time, numpy as np
import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
from common import *
from oracle import skirtref as sr
import skirt_b200 as sk
def check(name, S, n=20000):
    S.setup()
    t = S.grid_tables(); med = S.medium()
    e = sk.Engine(0); e.set_grid(t); e.medium(med['rho'], med['kext'], med['ksca'], med['g'])
    axes = (t['xv'],t['yv'],t['zv']) if t['kind']=='cartesian' else None
    r,k = rays(n, C1_BOX, 5)
    ra,ka = adversarial_rays(C1_BOX, axes)
    r = np.concatenate([r,ra]); k = np.concatenate([k,ka])
    t0=time.time(); a = S.path_batch(r,k,ell=0,nthreads=8); t1=time.time(); b = e.path_batch(r,k,ell=0); t2=time.time()
    same = paths_bit_identical(a,b)
    print(name, 'cells',S.Ncells,'segs',a['offsets'][-1], 'ref %.2fs gpu %.2fs'%(t1-t0,t2-t1), 'BIT-IDENTICAL' if same else 'DIFF', 'stuck', e.stuck_counts(), 'refwarn', sr.lib().skr_warnings())
    if not same:
        print(' counts equal', np.array_equal(a['offsets'],b['offsets']))
        if np.array_equal(a['offsets'],b['offsets']):
            print(' m equal', np.array_equal(a['m'],b['m']))
            for key in ('ds','s','dtau','tau'):
                d = np.abs(a[key]-b[key]); print(' ',key, d.max(), (d>0).sum())
        else:
            bad = np.nonzero(np.diff(a['offsets'])!=np.diff(b['offsets']))[0]; print(' bad rays', len(bad), bad[:10])
            i=bad[0]; print(r[i],k[i]); print(a['m'][a['offsets'][i]:a['offsets'][i+1]][:20]); print(b['m'][b['offsets'][i]:b['offsets'][i+1]][:20])
    tau_ref = S.opticaldepth_batch(r[:2000],k[:2000],0); tau_gpu = e.opticaldepth(r[:2000],k[:2000],0)
    print('  opticaldepth identical', np.array_equal(tau_ref,tau_gpu), ' whichcell identical', np.array_equal(S.whichcell(r[:5000]), e.whichcell(r[:5000])))
check('cart-lin', sr.RefSim(spec_c1(n=50), luminosities=[[1.0]], mixes=mix_v()))
check('cart-sympow', sr.RefSim(spec_c1(n=40, mesh='sympow 30'), luminosities=[[1.0]], mixes=mix_v()))
for s in (0,1,2):
    check('octtree-s%d'%s, sr.RefSim(spec_grid('octtree',search=s), luminosities=[[1.0]], mixes=mix_v()))
for s in (0,1):
    check('bintree-s%d'%s, sr.RefSim(spec_grid('bintree',search=s,maxlevel=12), luminosities=[[1.0]], mixes=mix_v()))
check('amesh', sr.RefSim(spec_grid('amesh'), luminosities=[[1.0]], mixes=mix_v(), amesh=make_amesh()))
check('voronoi', sr.RefSim(spec_grid('voronoi'), luminosities=[[1.0]], mixes=mix_v(), particles=voronoi_particles(5000)))
