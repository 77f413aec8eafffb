"""A/B aid: compare two dumps of tools/ab_dump.py (gpurun_out/ab_old.npz, gpurun_out/ab_new.npz)."""
import numpy as np
a=np.load("gpurun_out/ab_old.npz"); b=np.load("gpurun_out/ab_new.npz")
sa,sb=a["status"],b["status"]
d=sa!=sb
print("status differs", d.sum(), "of fitted", ((sa&28)>0).sum())
import collections
print(collections.Counter(zip(sa[d].tolist(), sb[d].tolist())).most_common(10))
print("by N:", collections.Counter(a["n"][d].tolist()))
same=~d & ((sa&12)>0)
rel=np.abs(a["chi2"][same]-b["chi2"][same])/np.abs(a["chi2"][same])
print("same-status converged: chi2 rel diff max %.3g, >1e-6: %d, >1e-3: %d"%(rel.max(), (rel>1e-6).sum(), (rel>1e-3).sum()))
