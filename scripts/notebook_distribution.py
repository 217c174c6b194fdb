import sys, json, numpy as np, time
sys.path.insert(0,'/root/repo/tests')
import oracle_lib as O
fx=json.load(open('/root/repo/tests/golden/notebook_fixtures.json'))["strong_vs_strong_1000_games"]
n=256; G=1000
ora=O.OracleBatch(n,mode=0,seed=5,n_threads=8)
# notebook: env constructed (reset(one_starts=True)), then env.reset() toggles -> first game one_starts False
ora.reset()
games=np.zeros(n,np.int64); obs_sum=np.zeros((n,18)); steps=np.zeros(n,np.int64); wdl=np.zeros((n,3)); rs=np.zeros((n,2))
t0=time.time(); tick=0
while (games<G).any():
    ro=ora.step(None,O.POL_STRONG,O.POL_STRONG,O.STEP_AUTORESET)
    live=games<G
    d=ro["done"].astype(bool)
    o=np.where(d[:,None],ro["final_obs"],ro["obs"]).astype(np.float64)
    obs_sum[live]+=o[live]; steps[live]+=1
    rs[live,0]+=ro["reward"][live]; rs[live,1]+=ro["reward2"][live]
    w=ro["info"][:,0]
    f=d&live
    wdl[f&(w==1),0]+=1; wdl[f&(w==0),1]+=1; wdl[f&(w==-1),2]+=1
    games+=f
    tick+=1
print("ticks",tick,"time",time.time()-t0)
means=obs_sum/steps[:,None]
def z(ref,s): return (ref-s.mean())/s.std(ddof=1)
for k in range(18):
    print("obs[%d] ref %.5f ours mean %.5f std %.5f z %.2f"%(k,fx["obs_mean"][k],means[:,k].mean(),means[:,k].std(ddof=1),z(fx["obs_mean"][k],means[:,k])))
print("steps ref",fx["total_steps"],"ours",steps.mean(),steps.std(ddof=1),z(fx["total_steps"],steps.astype(float)))
for i,nm in enumerate(["winners_plus1","winners_zero","winners_minus1"]):
    print(nm,fx[nm],wdl[:,i].mean(),wdl[:,i].std(ddof=1),z(fx[nm],wdl[:,i]))
for i in range(2):
    print("reward_sum",fx["reward_sums"][i],rs[:,i].mean(),rs[:,i].std(ddof=1),z(fx["reward_sums"][i],rs[:,i]))
np.savez('/tmp/nb_dist.npz',means=means,steps=steps,wdl=wdl,rs=rs)
