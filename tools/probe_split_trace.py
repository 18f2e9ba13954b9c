import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensorqtt_jl_b200 as q
q.default_context(0).set_profiling(True)
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe_split.py")).read())
