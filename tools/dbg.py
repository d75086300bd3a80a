import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sad_monte_carlo_b200 import WalkerEngine
eng = WalkerEngine(bench.lj31_config(64, lanes=1, flags=int(sys.argv[1]) if len(sys.argv) > 1 else 0))
eng.run(100)
print("ok", eng.walker(0).energy)
