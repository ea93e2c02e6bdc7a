"""Phase timing of one damped normal-equation step on the C4 synthetic block (GPU box)."""
import sys, os, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = subprocess.run([sys.executable, 'bench.py', '--steps', '5', '--warmup', '2'], capture_output=True, text=True,
                     cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b = json.loads(out.stdout.strip().splitlines()[-1])
print('ms/step %.3f' % b['ms_per_step'], {k: round(v, 3) for k, v in b['phases_ms_last_step'].items()})
