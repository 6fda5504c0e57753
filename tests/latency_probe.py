import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
import structured_latent_odes_b200 as slode
from structured_latent_odes_b200 import _cabi
dev="cuda"
torch.manual_seed(0)
m = slode.OdeModel(); m.init_with_params(torch.arange(0.,86.,device=dev),5,15,25,True,"midpoint",dev); m=m.to(dev)
z = torch.randn(128,15,device=dev); G = torch.randn(128,86,5,device=dev)
def wall(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e6
def fwd():
    with torch.no_grad(): m.solve_ODE(z)
def fwdbwd():
    m.zero_grad(set_to_none=True); m.solve_ODE(z).backward(G)
print("fwd us", wall(fwd)); print("fwd+bwd us", wall(fwdbwd))
# raw C call latency (fused fwd), no torch wrapper
L=_cabi.lib(); d=m.dynamics; n0,n2=m.latent_to_ode_net[0],m.latent_to_ode_net[2]
sol=torch.empty(86,128,5,device=dev); s=torch.cuda.current_stream().cuda_stream
args=(_cabi.METHOD_MIDPOINT,128,86,15,25,5,m.times.data_ptr(),z.data_ptr(),d.dynamics_hidden.weight.data_ptr(),d.dynamics_hidden.bias.data_ptr(),d.dyanamics_growth.weight.data_ptr(),d.dyanamics_growth.bias.data_ptr(),d.dyanmics_degradation.weight.data_ptr(),d.dyanmics_degradation.bias.data_ptr(),n0.weight.data_ptr(),n0.bias.data_ptr(),n2.weight.data_ptr(),n2.bias.data_ptr(),None,sol.data_ptr(),128*5,5,None,0,s)
print("raw C fwd call us", wall(lambda: L.slode_latent_fixed_fwd(*args)))
ev0,ev1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); ev0.record(); L.slode_latent_fixed_fwd(*args); ev1.record(); torch.cuda.synchronize(); print("device time of one fwd call us", ev0.elapsed_time(ev1)*1e3)
