#!/usr/bin/env python3
"""Work per ray class (camera / diffuse bounce / light-sample rays per light) of builder variants, counted on the CPU by
oracle/layout_walk.cpp over the product's own layout (no GPU): wide nodes visited, child boxes tested, leaves scanned.
usage: study_layout_work.py   (variants are environment settings read at layout build: TRT_COLLAPSE, TRT_REINSERT, ...)"""
import sys, os, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import oraclelib
import tinyraytracing_b200 as trt
from tinyraytracing_b200 import scenes, workloads
dn=os.open(os.devnull,os.O_WRONLY); sv=os.dup(1)
def shadow_rays(host, orc, n, rng):
    cam = workloads.camera_rays(host.camera(), n, rng)
    ids,t,pn,hp = orc.trace(cam, want_pn=True)
    ok = ids>=0; hp=hp[ok]; pn=pn[ok]
    # one diffuse bounce to get deeper vertices
    d = rng.normal(size=hp.shape); d/= np.linalg.norm(d,axis=1,keepdims=True); d[(d*pn).sum(1)<0]*=-1
    b = np.concatenate([hp,d],1).astype(np.float32)
    ids2,t2,pn2,hp2 = orc.trace(b, want_pn=True)
    pts = np.concatenate([hp, hp2[ids2>=0]])
    ls, lv, lvn, cum = host.lights()
    out={}
    for li,l in enumerate(ls):
        k = l["first_tri"] + rng.integers(0, max(1,min(l["n_tris"],24)), len(pts))
        bb = rng.random((len(pts),3)); bb/=bb.sum(1,keepdims=True)
        q = (lv[k].reshape(-1,3,3)*bb[:,:,None]).sum(1)
        dd = q-pts; dd/= np.linalg.norm(dd,axis=1,keepdims=True)
        out["shadow L%d"%li] = np.concatenate([pts,dd],1).astype(np.float32)
    return cam, b, out
for name,(w,h) in {"staircase":(96,54),"veach-mis":(96,54)}.items():
    with tempfile.TemporaryDirectory() as tmp:
        f=scenes.materialize(name,tmp,width=w,height=h)
        os.dup2(dn,1); host=trt.HostScene.load(f["xml"],f["obj"],f["mtl"],f["basedir"]); os.dup2(sv,1)
        orc=oraclelib.OracleScene(oraclelib.parsed_scene(name,w,h))
        rng=np.random.default_rng(3)
        cam,bounce,sh = shadow_rays(host,orc,30000,rng)
        sets={"camera":cam,"bounce":bounce}; sets.update(sh)
        for variant,env in {"greedy":{}, "optimal":{"TRT_COLLAPSE":"optimal"}, "noreinsert":{"TRT_REINSERT":"0"}}.items():
            for k in ("TRT_COLLAPSE","TRT_REINSERT"): os.environ.pop(k,None)
            os.environ.update(env)
            row=[]
            for sname,r in sets.items():
                ids,t,wk=oraclelib.walk_layout(host,r)
                row.append("%s n%.1f b%.1f l%.2f"%(sname[:9],wk["nodes"],wk["boxes"],wk["leaves"]))
            print(name, "%-10s"%variant, " | ".join(row))
