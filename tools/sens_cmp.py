import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
a = list(np.load(sys.argv[1], allow_pickle=True)); b = list(np.load(sys.argv[2], allow_pickle=True))
print(sys.argv[1], sys.argv[2], "fidelity defect", O.fidelity_defect(a, b), "linkdims equal", O.linkdims(a) == O.linkdims(b))
