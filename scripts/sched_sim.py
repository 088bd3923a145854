"""Event simulation of the attempt kernel schedule (148 SMs x 8 persistent warps x 32 lanes) with the measured pass time
per resident-warp count and the measured attempt histogram: base policy vs migrating thin warps in the tail, oracle
longest-first order, larger batches.  DESIGN.md 4.4."""
import heapq, random, sys
import numpy as np
random.seed(1); np.random.seed(1)
N = 65536; SMS = 148; WPS = 8
KS = np.array([1,2,3,4,5,6,7,8]); P = np.array([3,152,504,632,168,25,12,1], float); P /= P.sum()
TP = {0: 8.9, 1: 8.9, 2: 9.5, 3: 10.0, 4: 10.6, 5: 11.9, 6: 13.2, 7: 14.1, 8: 15.1}
def run(policy, thresh=12, seed=1):
    rng = np.random.RandomState(seed)
    ks = rng.choice(KS, size=N, p=P)
    cursor = 0
    resume = []          # list of remaining-attempt counts (migrated aircraft)
    warps = []           # per warp: list of remaining attempts per lane (0 = idle)
    live_on_sm = [WPS] * SMS
    ev = []
    lanes = [[0] * 32 for _ in range(SMS * WPS)]
    alive = [True] * (SMS * WPS)
    for w in range(SMS * WPS):
        heapq.heappush(ev, (0.0, w))
    t_end = 0.0; passes = 0; lane_att = 0
    active = SMS * WPS
    while ev:
        t, w = heapq.heappop(ev)
        sm = w // WPS
        L = lanes[w]
        # refill idle lanes
        for i in range(32):
            if L[i] == 0:
                if cursor < N:
                    L[i] = ks[cursor]; cursor += 1
                elif resume and policy != "base":
                    L[i] = resume.pop()
        nlive = sum(1 for x in L if x > 0)
        if nlive == 0:
            if cursor >= N and (policy == "base" or not resume):
                # exit (last-warp rule: if others are active they will take the list)
                alive[w] = False; live_on_sm[sm] -= 1; active -= 1
                if active == 0 and resume:
                    alive[w] = True; live_on_sm[sm] += 1; active += 1
                    heapq.heappush(ev, (t, w))
                t_end = max(t_end, t)
                continue
        if policy == "donate" and cursor >= N and 0 < nlive <= thresh and active > 2 * SMS:
            for i in range(32):
                if L[i] > 0: resume.append(L[i]); L[i] = 0
            alive[w] = False; live_on_sm[sm] -= 1; active -= 1
            t_end = max(t_end, t)
            continue
        dur = TP[live_on_sm[sm]]
        passes += 1; lane_att += nlive
        for i in range(32):
            if L[i] > 0: L[i] -= 1
        heapq.heappush(ev, (t + dur, w))
    # anything stranded?
    assert not resume, len(resume)
    return t_end, passes, lane_att / (32.0 * passes)
for pol, th in [("base", 0), ("donate", 8), ("donate", 12), ("donate", 16), ("donate", 20), ("donate", 24)]:
    r = [run(pol, th, s) for s in range(3)]
    print(pol, th, "makespan %.1f us  lane_eff %.3f" % (np.mean([x[0] for x in r]), np.mean([x[2] for x in r])))
# oracle LPT bound: same simulator, aircraft sorted by k descending
def run_sorted(seed=1):
    global KS
    rng = np.random.RandomState(seed)
    ks = np.sort(rng.choice(KS, size=N, p=P))[::-1]
    cursor = 0; ev = []; lanes = [[0]*32 for _ in range(SMS*WPS)]; live_on_sm=[WPS]*SMS
    for w in range(SMS*WPS): heapq.heappush(ev,(0.0,w))
    t_end=0
    while ev:
        t,w=heapq.heappop(ev); sm=w//WPS; L=lanes[w]
        for i in range(32):
            if L[i]==0 and cursor<N: L[i]=ks[cursor]; cursor+=1
        if not any(L): live_on_sm[sm]-=1; t_end=max(t_end,t); continue
        dur=TP[live_on_sm[sm]]
        for i in range(32):
            if L[i]>0: L[i]-=1
        heapq.heappush(ev,(t+dur,w))
    return t_end
print("oracle LPT", run_sorted())
# more envs per GPU: base policy at 2x and 4x the batch
for mult in (2, 4):
    N = 65536 * mult
    print("N x%d base makespan per 65536 envs: %.1f us" % (mult, run("base")[0] / mult))
