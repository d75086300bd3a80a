import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import bench
    from sad_monte_carlo_b200 import WalkerEngine
    cfg = bench.lj31_config(int(sys.argv[2]), lanes=1, walker_offset=int(sys.argv[1]))
    eng = WalkerEngine(cfg)
    print("ok", sys.argv[1:])
else:
    for off, W in [(0, 32), (0, 52), (0, 53), (52, 1), (52, 32), (1000, 64), (2000, 64), (3000, 256), (0, 1024)]:
        r = subprocess.run([sys.executable, __file__, str(off), str(W)], capture_output=True, text=True)
        print(off, W, "OK" if r.returncode == 0 else "FAIL " + r.stderr.strip().splitlines()[-1][:100], flush=True)
