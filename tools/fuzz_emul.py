"""Randomised API fuzzing of the kernel-logic emulator build against the oracle (distributed NTT with
random worker counts / pass limits / fused or copied exchange, whole-domain NTT, MSM with random
ranges, geometries and scalar distributions).  usage: python tools/fuzz_emul.py [seed] [seconds]"""
import sys, ctypes as C, numpy as np, random, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_plonk_b200._binding import bind, Context, DpError
from distributed_plonk_b200.worker import PlonkSlave
from distributed_plonk_b200 import dispatcher as disp
from oracle import loader as L
from tests import common
from tests.emul import build as _emul_build
lib=bind(C.CDLL(_emul_build.build(async_streams=os.environ.get('DP_TEST_EMUL_ASYNC','0')=='1')))  # DP_TEST_EMUL_ASYNC=1: real asynchronous streams
rng=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
def host_copy(d,s,n): C.memmove(d,s,n)
t0=time.time(); it=0; fails=0
while time.time()-t0 < float(sys.argv[2] if len(sys.argv)>2 else 60):
    it+=1
    kind=rng.choice(['fft','ntt','msm','poly','async'])
    try:
        if kind=='fft':
            W=rng.choice([1,1,2,4])
            ln=rng.randint(2 if W>1 else 0,8); lq=rng.randint(max(ln,4 if W==4 else 2),10)
            if W>1 and (1<<(ln>>1))<W: continue
            mc=rng.randint(1,11); ms=rng.randint(1,9)
            # plans must still be feasible: rows split in two needs each half <= limits
            ws=[PlonkSlave(lib,p,W) for p in range(W)]
            ok=True
            for w in ws:
                w.init([b""],1<<ln,1<<lq); w.ctx.debug_set_limits(mc,ms,0)
            if rng.random()<0.5 and W>1: common.attach_in_process(ws, 2*(1<<lq)*32//W)
            try:
                common.check_distributed_fft(L,ws,rng.choice([ln,lq]) if False else lq,True,rng.randint(1,1000),host_copy,n_in=rng.choice([None,(1<<lq)//8 or 1]))
            except DpError as e:
                if 'too large' not in str(e) and 'grid' not in str(e): raise
            for w in ws: w.close()
        elif kind=='ntt':
            c=Context(lib,0,0,1); c.init(np.zeros(0,dtype=np.uint8),1<<rng.randint(0,6),1<<rng.randint(0,9))
            mc=rng.randint(1,11); ms=rng.randint(1,9); c.debug_set_limits(mc,ms,0)
            ln=rng.randint(0,11)
            if ln>3*ms and ln>mc: c.close(); continue
            common.check_whole_ntt(L,c,ln,rng.randint(1,1000),n_in=rng.choice([None,rng.randint(0,1<<ln)]))
            c.close()
        elif kind=='poly':
            # rounds 3-5: random lengths around the 8 / 2048 chunk boundaries, random and special points,
            # quotient evaluations on random domain ratios
            ln=rng.randint(0,5); lq=ln+rng.randint(0,4)
            c=Context(lib,0,0,1); c.init(np.zeros(0,dtype=np.uint8),1<<ln,1<<lq)
            if rng.random()<0.4:
                common.check_quotient(L,c,1<<ln,1<<lq,rng.randint(1,10**6))
            n=rng.choice([1,2,7,8,9,63,64,65,2047,2048,2049,rng.randint(1,5000)])
            co=L.gen_fr(rng.randint(1,10**6),n)
            if rng.random()<0.3: co[rng.randrange(n):]=0
            pt=rng.choice([L.gen_fr(rng.randint(1,10**6),1)[0],np.zeros(4,dtype=np.uint64),common._fr_one(L),common._fr_neg_one(L)])
            ev=L.poly_eval(co,pt)
            assert np.array_equal(c.poly_eval(co,pt),ev),f"fuzz poly_eval n={n}"
            q,rem=c.poly_div_linear(co,pt)
            assert np.array_equal(rem,ev) and np.array_equal(q,L.poly_div_linear(co,pt)),f"fuzz poly_div n={n}"
            k=rng.randint(1,6); lens=[rng.choice([0,1,n,rng.randint(0,300)]) for _ in range(k)]
            if max(lens)==0: lens[0]=3
            polys=[L.gen_fr(rng.randint(1,10**6),ln_) for ln_ in lens]; cf=L.gen_fr(rng.randint(1,10**6),k)
            ol=rng.choice([None,rng.randint(1,max(lens)+5)])
            ref=L.poly_lincomb(polys,cf,ol if ol else None)
            assert np.array_equal(c.poly_lincomb(polys,cf,out_len=ol),ref),f"fuzz lincomb lens={lens} out_len={ol}"
            c.close()
        elif kind=='async':
            nb=rng.choice([5,33,300,600])
            bases=L.gen_bases(rng.randint(1,99),nb,min(nb,32),True)
            c=Context(lib,0,0,1); c.init(bases,1<<4,1<<6)
            c.debug_set_limits(11,9,rng.choice([0,0,4,7]))
            jobs={}
            for j in range(rng.randint(1,5)):
                lo=rng.randint(0,nb); hi=rng.randint(lo,nb); ns=rng.randint(0,hi-lo+3)
                sc=np.ascontiguousarray(common.scalar_sets(L,max(ns,1),rng.randint(1,999))[rng.choice(['uniform','witness-like','all one'])][:ns])
                c.msm_submit(j,lo,hi,sc); jobs[j]=(lo,hi,sc)
                if rng.random()<0.5:
                    x=L.gen_fr(rng.randint(1,999),16); assert np.array_equal(c.ntt(x,4,False,True),L.fft(x,False,True))
                if rng.random()<0.3 and jobs:
                    jj=rng.choice(list(jobs)); lo2,hi2,sc2=jobs.pop(jj); m2=min(hi2-lo2,sc2.shape[0])
                    common.assert_point_eq(L,c.msm_collect(jj),L.msm(bases[lo2:lo2+m2],sc2[:m2]),f"fuzz async msm [{lo2},{hi2})")
            order=list(jobs); rng.shuffle(order)
            for jj in order:
                lo2,hi2,sc2=jobs[jj]; m2=min(hi2-lo2,sc2.shape[0])
                common.assert_point_eq(L,c.msm_collect(jj),L.msm(bases[lo2:lo2+m2],sc2[:m2]),f"fuzz async msm [{lo2},{hi2})")
            c.close()
        else:
            nb=rng.choice([1,5,33,300,2048,2500])
            bases=L.gen_bases(rng.randint(1,99),nb,min(nb,32),True)
            W=rng.choice([1,1,2]); me=rng.randrange(W)
            c=Context(lib,0,me,W); c.init(bases,4,16)
            c.debug_set_limits(11,9,rng.choice([0,0,1,4,7,11]))
            for _ in range(2):
                lo=rng.randint(0,nb); hi=rng.randint(lo,nb); ns=rng.randint(0,hi-lo+3)
                name=rng.choice(['uniform','witness-like','all r-1','all zero','all one'])
                sc=common.scalar_sets(L,max(ns,1),rng.randint(1,999))[name][:ns]
                got=c.msm(lo,hi,sc)
                n=min(hi-lo,ns)
                common.assert_point_eq(L,got,L.msm(bases[lo:lo+n],sc[:n]),f"fuzz msm nb={nb} [{lo},{hi}) ns={ns} {name}")
            c.close()
    except AssertionError as e:
        fails+=1; print('FAIL',kind,str(e)[:200]); 
        if fails>3: break
print('iterations',it,'fails',fails)
sys.exit(1 if fails else 0)
