"""As sched_sim.py, for an order sorted on the attempt-count predictor (tests/tools/attempt_predictor.py)."""
import heapq, numpy as np
pairs = np.load("/tmp/kpred_pairs.npy")
N = 65536; SMS = 148; WPS = 8
TP = {0: 8.9, 1: 8.9, 2: 9.5, 3: 10.0, 4: 10.6, 5: 11.9, 6: 13.2, 7: 14.1, 8: 15.1}
def sim(ks):
    cursor = 0; ev = []; lanes = [[0]*32 for _ in range(SMS*WPS)]; live=[WPS]*SMS; n=len(ks)
    for w in range(SMS*WPS): heapq.heappush(ev,(0.0,w))
    t_end=0
    while ev:
        t,w=heapq.heappop(ev); sm=w//WPS; L=lanes[w]
        for i in range(32):
            if L[i]==0 and cursor<n: L[i]=ks[cursor]; cursor+=1
        if not any(L): live[sm]-=1; t_end=max(t_end,t); continue
        dur=TP[live[sm]]
        for i in range(32):
            if L[i]>0: L[i]-=1
        heapq.heappush(ev,(t+dur,w))
    return t_end
rng = np.random.RandomState(0)
idx = rng.randint(0, len(pairs), N)
pred, k = pairs[idx, 0], pairs[idx, 1]
print("base (random order, all attempts):", sim(list(k)))
rem = k - 1
first = 15.1 * 2   # first attempts in a natural-order kernel: 1.73 waves ~ 2 passes at 8 warps/SM
live = rem > 0
order = np.argsort(-np.minimum(pred[live], 8), kind="stable")
print("first attempt separately + pred-sorted remainder: %.1f + %.1f" % (first, sim(list(rem[live][order]))))
print("first attempt separately + random remainder: %.1f + %.1f" % (first, sim(list(rem[live]))))
order2 = np.argsort(-k[live], kind="stable")
print("first attempt separately + oracle-sorted remainder: %.1f + %.1f" % (first, sim(list(rem[live][order2]))))
