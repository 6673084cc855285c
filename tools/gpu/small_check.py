import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, bench, __graft_entry__ as ge
ge.build()
model,pool,params,ks=bench.make_workload("h2_4q",65536,0)
ev=model.evaluator(pool,3,0)
for i in range(3): out=ev.evaluate(ks,params,want_grad=True)
print("stats",ev.stats())
