import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, scenes
from positionbaseddynamics_b200 import _capi
from positionbaseddynamics_b200.model import HostModel
t0=time.perf_counter(); hm=HostModel(); scenes.cfg2(hm,1000,20); t1=time.perf_counter()
types,bodies,params,_=hm.constraints(); mass,_=hm.masses(); t2=time.perf_counter()
eng=_capi.Engine(0); eng.set_particles(hm.get("x"),mass); eng.add_flat(types,bodies,params); t3=time.perf_counter()
eng.color_first_fit(); t4=time.perf_counter()
eng.set_params(dt=0.005,sub_steps=1,max_iter=20); eng.step(1); eng.sync(); t5=time.perf_counter()
eng.step(1); eng.sync(); t6=time.perf_counter()
eng.set_mode(_capi.MODE_TILED); eng.step(1); eng.sync(); t7=time.perf_counter()
print("host model build %.2f s | export %.2f | engine add %.2f | colouring %.2f | first step (flatten+upload+capture) %.2f | second step %.4f | switch to tiled + step %.2f" % (t1-t0,t2-t1,t3-t2,t4-t3,t5-t4,t6-t5,t7-t6))
