"""Discrete simulation of the barrier protocol of tc_scan_kernel (csrc/tcscan.cu): producer, two MMA issuers (one per
accumulator stage), four epilogue groups taking the 128-column halves of the N-tiles in turn; mbarriers with phase
parity and arrival counts (an arrival on a completed phase is an error), asynchronous arrivals (bulk copies, tcgen05.commit)
delivered at random times.  Finds protocol deadlocks on the CPU: a first version signalled "accumulator ready" per stage
only and hung, because a parity wait cannot tell phase k from phase k + 2 when a group skips a phase."""
import random, sys
class MBar:
    def __init__(s, count, name): s.count=count; s.pending=count; s.phase=0; s.name=name
    def arrive(s):
        assert s.pending>0, f"arrive underflow on {s.name}"
        s.pending-=1
        if s.pending==0: s.phase^=1; s.pending=s.count
    def done(s, parity): return s.phase != parity   # try_wait.parity(p): true when phase with parity p has completed
def run(n_mtiles, bns, stages=16, seed=0, async_delay=True):
    rnd=random.Random(seed)
    n_tiles=len(bns)
    afull=[MBar(1,f'afull{i}') for i in range(stages)]; aempty=[MBar(2,f'aempty{i}') for i in range(stages)]
    tfull=[[MBar(1,f'tfull_g{i}_s{j}') for j in range(2)] for i in range(4)]; tempty=[MBar(8,f'tempty{i}') for i in range(2)]
    pending=[]   # async arrivals (commit / bulk copy): (bar)
    def producer():
        stage=0; phase=0
        for t in range(n_mtiles):
            while not aempty[stage].done(phase^1): yield
            pending.append(afull[stage])
            stage+=1
            if stage==stages: stage=0; phase^=1
    def issuer(me):
        stage=0; phase=0; nseq=0; acc_phase=0; iseq=0
        for t in range(n_mtiles):
            while not afull[stage].done(phase): yield
            for nt in range(n_tiles):
                n_halves=2 if (bns[nt]>>5)>4 else 1
                if (nseq&1)==me:
                    while not tempty[me].done(acc_phase^1): yield
                    acc_phase^=1
                    for hf in range(n_halves): pending.append(tfull[(iseq+hf)&3][me])
                iseq+=n_halves
                nseq+=1
            pending.append(aempty[stage])
            stage+=1
            if stage==stages: stage=0; phase^=1
            yield
    def epi(gidx, w):
        nseq=0; iseq=0; mine=[0,0]
        for t in range(n_mtiles):
            for nt in range(n_tiles):
                n_chunks=bns[nt]>>5; n_halves=2 if n_chunks>4 else 1
                for hf in range(n_halves):
                    if (iseq&3)==gidx:
                        acc=nseq&1
                        while not tfull[gidx][acc].done(mine[acc]&1): yield
                        mine[acc]+=1
                        yield
                        tempty[acc].arrive()
                        if n_halves==1: tempty[acc].arrive()
                        yield
                    iseq+=1
                nseq+=1
    procs=[producer(), issuer(0), issuer(1)]+[epi(g,w) for g in range(4) for w in range(4)]
    alive=list(procs); steps=0
    while alive:
        steps+=1
        if steps>5_000_000: print("DEADLOCK/livelock", [(b.name,b.phase,b.pending) for b in afull[:2]+aempty[:2]+tempty]); return False
        if pending and (not async_delay or rnd.random()<0.5):
            b=pending.pop(rnd.randrange(len(pending)) if False else 0); b.arrive(); continue
        p=rnd.choice(alive)
        try: next(p)
        except StopIteration: alive.remove(p)
    while pending: pending.pop(0).arrive()
    return True

def run4i(n_mtiles, n_tiles, split=1, stages=16, seed=0, n_release=None, max_steps=5_000_000):
    """The protocol of tc_scan4i_kernel: stage = issuer = epilogue group; N-tile number s = mt * n_tiles + nt of the CTA's
    running sequence belongs to pipeline s % 4; an issuer only waits for / releases (tcgen05.commit on aempty) the stream
    tiles it has an N-tile of, so aempty counts min(n_tiles, 4) arrivals; split = 2: two half-stages per pipeline, used in
    turn.  Returns False on a deadlock; asserts on an arrival that would complete a phase twice."""
    rnd = random.Random(seed)
    n_release = min(n_tiles, 4) if n_release is None else n_release
    afull = [MBar(1, f'afull{i}') for i in range(stages)]; aempty = [MBar(n_release, f'aempty{i}') for i in range(stages)]
    tfull = [[MBar(1, f'tfull{i}.{h}') for h in range(split)] for i in range(4)]
    tempty = [[MBar(4, f'tempty{i}.{h}') for h in range(split)] for i in range(4)]
    pending = []
    done_items = [0] * 4
    def items(me):
        mt, nt = 0, me
        while nt >= n_tiles: nt -= n_tiles; mt += 1
        while mt < n_mtiles:
            mt2, nt2 = mt, nt + 4
            while nt2 >= n_tiles: nt2 -= n_tiles; mt2 += 1
            yield mt, nt, mt2
            mt, nt = mt2, nt2
    def producer():
        stage = 0; phase = 0
        for t in range(n_mtiles):
            while not aempty[stage].done(phase ^ 1): yield
            pending.append(afull[stage])
            stage += 1
            if stage == stages: stage = 0; phase ^= 1
    def issuer(me):
        use = 0; ready = -1
        for mt, nt, mt2 in items(me):
            stage = mt % stages
            if mt != ready:
                while not afull[stage].done((mt // stages) & 1): yield
                ready = mt
            hs = use % split
            while not tempty[me][hs].done(((use // split) & 1) ^ 1): yield
            use += 1
            pending.append(tfull[me][hs])
            if mt2 != mt: pending.append(aempty[stage])
            yield
    def epi(g, w):
        use = 0
        for mt, nt, mt2 in items(g):
            hs = use % split
            while not tfull[g][hs].done((use // split) & 1): yield
            use += 1
            yield
            tempty[g][hs].arrive()
            if w == 0: done_items[g] += 1
            yield
    procs = [producer()] + [issuer(i) for i in range(4)] + [epi(g, w) for g in range(4) for w in range(4)]
    alive = list(procs); steps = 0
    while alive:
        steps += 1
        if steps > max_steps: return False
        if pending and rnd.random() < 0.5:
            pending.pop(rnd.randrange(len(pending))).arrive(); continue
        p = rnd.choice(alive)
        try: next(p)
        except StopIteration: alive.remove(p)
    while pending: pending.pop(0).arrive()
    return sum(done_items) == n_mtiles * n_tiles
if __name__ == "__main__":
    rnd = random.Random(5)
    cases = [[256, 96], [64], [256], [256, 256, 128], [96, 64, 32], [256] * 8, [160, 256, 32, 224, 128]]
    cases += [[32 * rnd.randint(1, 8) for _ in range(rnd.randint(1, 8))] for _ in range(12)]
    bad = 0
    for bns in cases:
        for seed in range(3):
            if not run(60, bns, seed=seed):
                print("FAILED", bns, seed); bad += 1
    print("cases", len(cases), "failures", bad)
    bad4 = sum(not run4i(50, nt, split=sp, seed=sd) for nt in range(1, 17) for sp in (1, 2) for sd in range(2))
    print("tc_scan4i protocol: failures", bad4)
