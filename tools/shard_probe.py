import sys,os; sys.path.insert(0,".")
import __graft_entry__ as e; e.load_package()
from graph_embed_b200 import capi, graphs, sharding
ctx=capi.Context(0)
A=graphs.rgg(500000,10.0,seed=7); n=A.shape[0]
x0=capi.reference_uniform(23,n*2).reshape(n,2)
for world in (1,2,4,8):
    r0,r1,R,ld=sharding.row_block(n,world,0)
    plan=ctx.flat_plan(A,2,capi.flat_params(),rows=(r0,r1))
    plan.upload(x0); plan.iterate(2); plan.sync(); plan.profile(True); plan.iterate(3)
    p=plan.profile_get(); ms=p["repulsion_ms"]/p["repulsion_launches"]
    print("world",world,"rows",r1-r0,"rep ms %.2f"%ms,"ideal %.2f"%(215.4/world), flush=True)
    plan.close()
