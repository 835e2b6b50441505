"""Ray-sample balance of sort-first partitions, off line: per-pixel ray-sample counts of the full-size cfg3 / cfg4 frames from the CPU oracle
(a tiny volume and 0 LIC steps: the counts depend on the geometry only), summed per 16x16 block and dealt to 2 / 4 / 8 ranks in units of
ux x uy blocks with every row rotation (skew) 0..8: imbalance = busiest rank / mean - 1.  python scripts/partition_balance_sim.py
(test infrastructure: imports oracle/)"""
import sys, numpy as np
sys.path.insert(0,'/root/repo')
from vectorvisualization_b200 import configs, fields as F
from oracle import vvo
import copy

def counts(cfgname, camera=None):
    mk = getattr(configs, cfgname)
    kw = dict(n=8)
    if cfgname == 'cfg4': kw['noise_n'] = 8
    s = mk(**kw) if camera is None else mk(camera=camera, **kw)
    full = getattr(configs, cfgname)
    # full-size image and step, tiny volume (counts depend on geometry only)
    s.width = s.height = {'cfg3':1024,'cfg4':2048}[cfgname]
    p = dict(s.params); p['stepsForward']=0; p['stepsBackward']=0; s.params=p
    o = vvo.OracleScene(s)
    img, cnt, tot = o.raycast()
    return cnt, tot

def block_sums(cnt):
    h,w = cnt.shape
    nby, nbx = (h+15)//16, (w+15)//16
    pad = np.zeros((nby*16, nbx*16), np.int64); pad[:h,:w]=cnt
    return pad.reshape(nby,16,nbx,16).sum(axis=(1,3))

def imbalance(bs, world, ux, uy, skew):
    nby, nbx = bs.shape
    nux, nuy = (nbx+ux-1)//ux, (nby+uy-1)//uy
    tot = np.zeros(world)
    for by in range(nby):
        for bx in range(nbx):
            uxi, uyi = bx//ux, by//uy
            uid = uyi*nux + (uxi + skew*uyi) % nux
            tot[uid % world] += bs[by,bx]
    return tot.max()/tot.mean()-1

if __name__ == '__main__':
    scenes = [('cfg3',None),('cfg3',dict(F.CAMERA_CLOSE)),('cfg4',None)]
    res = {}
    for name,cam in scenes:
        cnt,tot = counts(name,cam)
        print(name, 'close' if cam else 'default', tot, flush=True)
        bs = block_sums(cnt)
        for world in (2,4,8):
            for (ux,uy) in ((1,1),(2,1),(1,2),(2,2),(4,2),(4,4)):
                best = []
                for skew in range(0, 9):
                    best.append((imbalance(bs, world, ux, uy, skew), skew))
                res[(name, bool(cam), world, ux, uy)] = best
                b = sorted(best)[:3]
                print('  world %d unit %dx%d: ' % (world,ux,uy) + ', '.join('skew %d: %.2f%%' % (s, 100*i) for i,s in b) + '   | skew 1: %.2f%% skew 3: %.2f%%' % (100*best[1][0], 100*best[3][0]), flush=True)
