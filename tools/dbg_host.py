import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from signalops import Amplify, Filt, Lowpass, Signal, GPUSink, cabi, dB, kHz
from signalops.lowering import lower
def chain(x): return Signal(x, 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(-20 * dB)
gpu = GPUSink([0])
rng = np.random.default_rng(1)
n, nch, ninst = 48000, 2, 40
xs = [np.asfortranarray(rng.standard_normal((n, nch))) for _ in range(ninst)]
cp = gpu.compiled(lower(chain(xs[0])).tobytes())
ys = [np.zeros((n, nch), order="F") for _ in range(ninst)]
for name, fn in (("pageable", lambda: cp.run_host(ninst, xs, ys)),):
    try:
        print(name, fn())
    except Exception as e:
        print(name, "FAILED", e)
xp = [cabi.pinned_empty((n, nch), np.float64) for _ in range(ninst)]
for a, b in zip(xp, xs): a[...] = b
yp = [cabi.pinned_empty((n, nch), np.float64) for _ in range(ninst)]
try: print("pinned", cp.run_host(ninst, xp, yp))
except Exception as e: print("pinned FAILED", e)
try: print("mixed", cp.run_host(ninst, xp, ys))
except Exception as e: print("mixed FAILED", e)
try: print("one", cp.run_host(1, xs[:1], ys[:1]))
except Exception as e: print("one FAILED", e)
