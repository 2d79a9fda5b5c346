import sys, time, os
sys.argv=['bench.py']; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import numpy as np, torch, bench
import __graft_entry__ as ge
pkg=ge.load_package(); L=sys.modules[pkg.__name__+'._lib']
w=bench.WORKLOADS['c3']; N,T,W,H,M,B=w['N'],w['T'],w['W'],w['H'],w['M'],w['B']
import synthdata
cam_K=synthdata.camera_for(W,H)
torch.cuda.init(); torch.zeros(1,device='cuda')
for rep in range(2):
    t0=time.perf_counter()
    opt=pkg.SMPLDepthSequenceOptimizer(image_size=(W,H),num_frames=T,cam_K=cam_K,device='cuda:0',smpl_model_parameters_path=bench.model_dir(),scene_update=False,max_scene_points=M,**bench.COEFS)
    t1=time.perf_counter()
    opt._make_context(T,N,B)
    torch.cuda.synchronize(); t2=time.perf_counter()
    print(f'construct {t1-t0:.3f} s | context {t2-t1:.3f} s')
    import cProfile,pstats
    opt.ctx.close()
    pr=cProfile.Profile(); pr.enable(); opt._make_context(T,N,B); torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(8)
    opt.ctx.close()
