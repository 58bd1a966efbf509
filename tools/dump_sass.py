"""Developer tool: compile a bench config's kernel and dump SASS with line info.
usage: python tools/dump_sass.py c2_skin [out.sass]"""
import sys, os, importlib, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import benchcfg
name = sys.argv[1]
mc = importlib.import_module('pyxopto_b200.%s.mc' % benchcfg.GEOMETRY[name])
sim = benchcfg.CONFIGS[name](mc)
block = int(sys.argv[3]) if len(sys.argv) > 3 else 256
cubin, log, hit = sim.compile(1000, block=block)
path = '/tmp/%s.cubin' % name
open(path, 'wb').write(cubin)
out = sys.argv[2] if len(sys.argv) > 2 else '/tmp/%s.sass' % name
dis = subprocess.run(['nvdisasm', '-g', '-c', path], capture_output=True, text=True).stdout
open(out, 'w').write(dis)
n = sum(1 for l in dis.splitlines() if l.strip().startswith('/*') and ';' in l)
print(path, out, 'instructions:', n)
subprocess.run('cuobjdump -res-usage %s | tail -3' % path, shell=True)
