import sys, ctypes as C
import numpy as np
sys.path.insert(0,'.')
from minicv_b200 import estimation as est
g=np.load('tests/golden/ransac_e.npz')
a,b,Ec,n=g["min5_a"][0],g["min5_b"][0],g["min5_E"][0],g["min5_n"][0]
Es=est.fivepoint(a,b)
print('cv2 n',n)
for i in range(n): print(' cv2 E0=%.15g E8=%.15g'%(Ec[3*i,0],Ec[3*i+2,2]))
for e in Es: print(' lib E0=%.15g E8=%.15g  fro=%.6f  det=%.3e  2EEtE-tr=%.3e'%(e[0,0],e[2,2],np.sqrt((e*e).sum()),np.linalg.det(e),np.abs(2*e@e.T@e-np.trace(e@e.T)*e).max()))
