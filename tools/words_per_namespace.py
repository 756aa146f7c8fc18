"""How many 32-throttle words a pod visits (what a pass costs per pod) under different orders of the throttle columns, for the
synthetic BASELINE configs: the generator's order, a random order (the order objects happen to arrive in a real cluster) and
the order the host layer lays its columns out in (kt_host.cc reorder_columns: Throttles by namespace, ClusterThrottles by the
set of namespaces their namespaceSelectors admit).  CPU only.      python tools/words_per_namespace.py [C2 C3 ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kube_throttler_b200 import synth, abi
def analyse(cfg):
    snap = synth.generate(cfg)
    m=snap.m; NS=snap.n_ns
    nl = snap.ns_labels
    nsl=[]
    for ns in range(NS):
        d={}
        for s in range(nl.shape[0]):
            v=int(nl[s,ns])
            if v!=abi.LABEL_EMPTY: d[v>>32]=v&0xffffffff
        nsl.append(d)
    term_off=snap.term_off; ns_off=snap.ns_req_off
    def term_ns_ok(term, ns):
        for r in range(ns_off[term],ns_off[term+1]):
            k=int(snap.req_key[r]); op=int(snap.req_op[r]); vals=[int(x) for x in snap.req_vals[snap.req_val_off[r]:snap.req_val_off[r+1]]]
            has=k in nsl[ns]; v=nsl[ns].get(k)
            ok = (has and v in vals) if op==abi.OP_IN else ((not has) or v not in vals) if op==abi.OP_NOTIN else has if op==abi.OP_EXISTS else (not has)
            if not ok: return False
        return True
    applies=np.zeros((m,NS),bool)
    for t in range(m):
        if snap.kind[t]==abi.KIND_THROTTLE: applies[t,int(snap.thr_ns[t])]=True
        else:
            for term in range(term_off[t],term_off[t+1]):
                for ns in range(NS):
                    if term_ns_ok(term,ns): applies[t,ns]=True
    cnt=np.bincount(snap.running.ns_id, minlength=NS)
    def words(order):
        pos=np.empty(m,int); pos[order]=np.arange(m); w=pos//32
        wp=np.array([len(set(w[applies[:,ns]])) for ns in range(NS)])
        return wp.mean(), wp.max(), (wp*cnt).sum()/cnt.sum()
    rng=np.random.default_rng(0)
    orders={"synth order":np.arange(m), "random creation order":rng.permutation(m),
            "sorted by (kind, namespace set)":np.array(sorted(range(m),key=lambda t:(int(snap.kind[t]), tuple(applies[t].tolist()), t)))}
    print(cfg)
    for k,o in orders.items(): print("  %-36s words per namespace mean %.1f max %d | per running pod %.2f" % ((k,)+words(o)))
for cfg in (sys.argv[1:] or ["C2", "C3"]):
    analyse(cfg)
