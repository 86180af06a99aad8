"""Run bench.py against an experimental build of the library:
  python tools/exp/run_variant.py tools/exp/lib896.so [bench.py arguments]"""
import os, runpy, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
from dragnet_b200 import native
native.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = [os.path.join(R, 'bench.py')] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name='__main__')
