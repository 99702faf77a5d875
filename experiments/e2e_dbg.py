import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from speaker_recognition_b200 import synth
from speaker_recognition_b200.gmm.pygmm import GMM
from speaker_recognition_b200.gmmset import GMMSet
from speaker_recognition_b200.pipeline import BatchScorer
S,K,D=100,128,39
W,M,Sg=synth.models(S,K,D,seed0=1000)
gs=GMMSet(gmm_order=K)
for i in range(S):
    gs.gmms.append(GMM(K).set_params(W[i],M[i],Sg[i])); gs.y.append(i)
B=64; U=2560768
base=[synth.audio(U,seed=u) for u in range(4)]
host=torch.empty(B*U,dtype=torch.int16).pin_memory()
hv=host.numpy()
for u in range(B): hv[u*U:(u+1)*U]=base[u%4]
lengths=[U]*B
bs=BatchScorer(gs,fs=16000,nd=2)
for _ in range(3): bs.predict_pcm(host,lengths)
torch.cuda.synchronize()
log=[]
orig=bs.submit
def sub(p,l):
    t=orig(p,l); log.append(len(bs.chunks)); return t
bs.submit=sub
for mode in ("auto","pipe2"):
    if mode=="pipe2":
        bs.chunk_sets["sync"]=bs.chunk_sets["pipe"]
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    t0=time.perf_counter()
    pending=bs.submit(host,lengths)
    for _ in range(19):
        nxt=bs.submit(host,lengths); bs.result(pending); pending=nxt
    bs.result(pending)
    e1.record(); torch.cuda.synchronize()
    print(mode, e0.elapsed_time(e1)/20, log[-20:])
