import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import esac_b200.api as api
from oracle import esac_oracle as O
z = np.load('/root/repo/tests/golden/world_scale_24x32.npz')
p = z['params'].tolist(); params = (int(p[0]), int(p[1])) + tuple(float(v) for v in p[2:9]) + (int(p[9]),)
coords, assign = z['coords'], z['assign']
g_ref = np.zeros_like(coords)
loss_ref, bt = O.backward(coords, g_ref, assign, z['gt_pose'], 1.0, 100.0, 100.0, *params, seed=int(z['seed']), trace=True)
api.context().set_option('fixed_seed', 1)
api.set_seed(int(z['seed']))
g = np.zeros_like(coords)
loss = api.backward(coords, g, assign, z['gt_pose'], 1.0, 100.0, 100.0, *params)
hy = api.last_hypotheses(losses=True)
print('loss gpu', loss, 'ref', loss_ref)
print('probs diff', np.abs(hy['probs']-bt.probs).max())
for h in range(len(assign)):
    if bt.probs[h] < 1e-3: continue
    r = np.concatenate([bt.ref[h][0].ravel(), bt.ref[h][1].ravel()])
    print(h, 'p %.4f'%bt.probs[h], 'loss ref %.6f gpu %.6f'%(bt.losses[h], hy['losses'][h]), 'pose diff', np.abs(r-hy['refined'][h]).max(), 'init diff', np.abs(np.concatenate([bt.hyps[h].rvec.ravel(), bt.hyps[h].tvec.ravel()])-hy['poses'][h]).max())
print('gt pose', z['gt_pose'])
