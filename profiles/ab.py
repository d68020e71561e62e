"""python profiles/ab.py [rounds] name=lib.so[@ENV=VAL[,ENV=VAL]] ...  -- interleaved A/B of library builds on one box."""
import os, subprocess, sys
args = sys.argv[1:]
rounds = int(args.pop(0)) if args and args[0].isdigit() else 2
for rnd in range(rounds):
    for a in args:
        name, path = a.split('=', 1)
        extra = {}
        if '@' in path:
            path, envs = path.split('@', 1)
            extra = dict(e.split('=', 1) for e in envs.split(','))
        env = dict(os.environ, PBR_B200_LIB=os.path.abspath(path), **extra)
        out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), 'ab_one.py'), name], env=env,
                             capture_output=True, text=True, timeout=600)
        lines = [l for l in out.stdout.splitlines() if l.startswith('AB ')]
        print(lines[-1] if lines else f"AB {name}: FAILED rc={out.returncode} {out.stderr[-1500:]}", flush=True)
