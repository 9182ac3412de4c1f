import sys; sys.path.insert(0,'tests')
import numpy as np, torch
from helpers import *
env_id,N,G,area,n_obs='LinearDrone',10,2,1.5,4
agent, goal, obs = random_scene(env_id, N, G, area, n_obs, 1)
env = product_env(env_id, N, area, n_obs); env.edge_cap_per_agent=64
pobs = product_obstacles(env_id, obs)
graph = env.get_graph(torch.from_numpy(agent).cuda(), torch.from_numpy(goal).cuda(), pobs)
hits = graph.hits.cpu().numpy()
oenv = oracle_env(env_id, N, area, n_obs)
packed = pobs.packed.cpu().numpy()
from oracle.geometry import raytracing
for g in range(G):
    oobs = oracle_obstacles(packed[g])
    pos = torch.from_numpy(agent[g][:, :3])
    starts = pos[:, None, :].expand(N, oenv.ray_table.shape[0], 3).contiguous()
    hs, al, order = raytracing(starts, starts + oenv.ray_table, oobs, 16)
    bad = np.argwhere(hits[g] != hs.numpy())
    print('graph', g, 'n mismatch', len(bad))
    for i in sorted(set(bad[:,0].tolist()))[:3]:
        print(' agent', i, 'order', order[i].tolist(), 'alphas', al[i].tolist())
        print('  got ', hits[g][i][:4].tolist())
        print('  want', hs[i][:4].tolist())
        # find which rays product picked
        allh = (starts[i] + (starts[i]+oenv.ray_table - starts[i]) * 1.0)
print('tab equal', np.array_equal(env._ray_table_np, oenv.ray_table.numpy()))
g=1
oobs = oracle_obstacles(packed[g])
pos = torch.from_numpy(agent[g][:, :3])
starts = pos[:, None, :].expand(N, oenv.ray_table.shape[0], 3).contiguous()
hs, al, order = raytracing(starts, starts + oenv.ray_table, oobs, 16)
bad = np.argwhere(hits[g] != hs.numpy())
print(bad)
for b in bad:
    i,k,c = b
    print('rank',k,'comp',c,'got',repr(hits[g][i,k,c]),'want',repr(hs[i,k,c].item()), 'ray', order[i,k].item(), 'alpha', repr(al[i,k].item()))
    print(' got hit', hits[g][i,k], 'want', hs[i,k].numpy())
    r = order[i,k].item()
    print(' table', oenv.ray_table[r].numpy(), 'start', starts[i,r].numpy())
