"""debug helper: one parity case outside pytest (library error messages reach the terminal)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fftw3_b200 import binding as B
import fftcheck as F
lib = B.load()
kind, prec, n, hm, ip = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
fn = {"r2c": F.r2c, "c2r": F.c2r, "c2c": F.c2c}[kind]
print(kind, prec, n, hm, ip, fn(lib, prec, (n,), howmany=hm, inplace=bool(ip)), flush=True)
