import sys, numpy as np, ctypes
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import oracle
from irbaboon_b200 import engine as eng, synth
orc=oracle.Oracle()
x = synth.white_noise(1001, 0, 3000); h = synth.decaying_ir(2000, 900)
want=orc.convolve_nonperiodic(x,h)
got=eng.convolve_nonperiodic(x,h)
print("py api", np.abs(got-want).max(), np.abs(want).max())
got=eng.convolve_nonperiodic(x,h)
print("py api 2nd", np.abs(got-want).max())
x2=synth.white_noise(1001,0,1500); h2=synth.decaying_ir(2000,600)
print("1500/600", np.abs(eng.convolve_nonperiodic(x2,h2)-orc.convolve_nonperiodic(x2,h2)).max())
got=eng.convolve_nonperiodic(x,h)
print("py api 3rd", np.abs(got-want).max(), np.isnan(got).sum(), np.abs(got).max())
d=np.abs(got-want)[0]; print(np.argmax(d), d[:5], d[2990:3010], d[-5:])
