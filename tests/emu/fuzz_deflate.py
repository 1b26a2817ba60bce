"""Randomized run of the emulated deflate kernel at random levels with and without the one-stream flag (MZ_CUDA_FLAG_DICT) over runs,
short and long periods, text, records and mixtures; every stream must inflate to its input with zlib.
  python tests/emu/fuzz_deflate.py <seed> <seconds>      (needs tests/emu/libmzemu.so)   TEST INFRASTRUCTURE ONLY."""
import os, sys, zlib, random, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emushim, datagen
emu = emushim.EmuLib()
rng = random.Random(int(sys.argv[1])); T=float(sys.argv[2])
def gen(n):
    k = rng.randrange(7)
    if k==0: return datagen.text_like(n, rng.randrange(1<<20))
    if k==1: return datagen.mixed(n, rng.randrange(1<<20))
    if k==2: return datagen.binary_records(n, rng.randrange(1<<20))
    if k==3: return bytes([rng.randrange(256)])*n
    if k==4:
        p = bytes(rng.randrange(256) for _ in range(rng.randrange(1,9))); return (p*(n//len(p)+1))[:n]
    if k==5:
        p = datagen.random_bytes(rng.randrange(100, 50000), rng.randrange(1<<20)); return (p*(n//len(p)+1))[:n]
    parts=[]; left=n
    while left>0:
        m=min(left, rng.randrange(1,30000)); parts.append(gen(m) if rng.random()<0.7 else bytes([rng.randrange(3)])*m); left-=m
    return b"".join(parts)
t0=time.time(); it=0; fails=0
while time.time()-t0 < T:
    it+=1
    n = rng.choice([32768,32769,65535,65536,65537,98304,131072,131073,200000,300001]) if rng.random()<0.5 else rng.randrange(0,400000)
    data = gen(n)
    level = rng.choice([1,3,5,6,7,9]); flags = rng.choice([1,3,3])
    try:
        comp,_ = emu.deflate(data, level=level, final=flags)
        if zlib.decompress(comp,-15)!=data: raise Exception("deflate mismatch")
    except Exception as e:
        fails+=1; print("DEFLATE FAIL", it, n, level, flags, repr(e)[:100]); open('/tmp/fuzz_deflate_fail_%d.bin'%it,'wb').write(data)
print("iterations", it, "fails", fails)
