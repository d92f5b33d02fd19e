"""Where the wall time of examples/example1.py goes (cProfile, cumulative) - small-problem regime."""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'examples'))
import torch
import example1
example1.run(refine_time=3, epsilon_range=1, element_type='P2', out=lambda *a: None)   # warm-up: library load, first launches
torch.cuda.synchronize()
t = time.time()
pr = cProfile.Profile(); pr.enable()
example1.run(refine_time=6, epsilon_range=6, element_type='P2', out=lambda *a: None)
pr.disable()
print('example1.run(6, 6, P2) seconds:', time.time() - t)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:9000])
