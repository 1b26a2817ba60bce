"""Randomized differential run of the emulated kernels against zlib: deflate at every level, inflate of zlib streams with
random levels / strategies / flushes through random windows, K6 rounds with random segment sizes.
  python tests/emu/fuzz_kernels.py <seed> <seconds>      (needs tests/emu/libmzemu.so)   TEST INFRASTRUCTURE ONLY."""
import os, sys, zlib, random, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emushim, datagen
emu = emushim.EmuLib()
rng = random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
T = float(sys.argv[2]) if len(sys.argv)>2 else 300
def gen(n):
    k = rng.randrange(8)
    if k==0: return datagen.text_like(n, rng.randrange(1<<20))
    if k==1: return datagen.mixed(n, rng.randrange(1<<20))
    if k==2: return datagen.random_bytes(n, rng.randrange(1<<20))
    if k==3: return datagen.binary_records(n, rng.randrange(1<<20))
    if k==4: return bytes([rng.randrange(256)])*n
    if k==5:
        p = bytes(rng.randrange(256) for _ in range(rng.randrange(1,300))); return (p*(n//len(p)+1))[:n]
    if k==6:
        a = rng.randrange(2,40); return bytes(rng.randrange(a) for _ in range(n))
    parts=[]; left=n
    while left>0:
        m=min(left, rng.randrange(1,20000)); parts.append(gen(m) if rng.random()<0.9 else bytes(m)); left-=m
    return b"".join(parts)
t0=time.time(); it=0; fails=0
while time.time()-t0 < T:
    it+=1
    n = rng.choice([0,1,2,3,5,31,32,33,255,258,259,1000,4095,4096,32767,32768,32769,65535,65536,65537,100000,140001,200000]) if rng.random()<0.5 else rng.randrange(0,220000)
    data = gen(n)
    # deflate
    level = rng.choice([0,1,2,3,4,5,6,7,8,9])
    try:
        comp,_ = emu.deflate(data, level=level, final=True)
        if zlib.decompress(comp,-15)!=data: raise Exception("deflate mismatch")
    except Exception as e:
        fails+=1; print("DEFLATE FAIL", it, n, level, repr(e)[:100]); open('/tmp/fuzz_fail_%d.bin'%it,'wb').write(data)
    # zlib stream with random flushes
    zl = rng.choice([0,1,3,6,9]); strat = rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED])
    co = zlib.compressobj(zl, zlib.DEFLATED, -15, rng.choice([1,8,9]), strat)
    parts=[]; o=0
    while o < len(data):
        m = rng.randrange(1, 60000); parts.append(co.compress(data[o:o+m])); o+=m
        if rng.random()<0.3: parts.append(co.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])))
    parts.append(co.flush()); z=b"".join(parts)
    try:
        iw = rng.choice([0,0,0,100,3000,70000]); ow = rng.choice([0,0,0,300,5000,66000])
        st,out,cons,_ = emu.inflate(z + bytes(rng.randrange(0,9)), len(data), iw, ow)
        if not (st==1 and out==data and cons==len(z)): raise Exception("inflate st %d len %d cons %d/%d"%(st,len(out),cons,len(z)))
        if len(z) > 3000:
            seg = rng.choice([512,1024,2048,4096,8192]); win = rng.choice([0,0,16384,65536])
            st,out,cons,stats = emu.inflate_spec(z, len(data), seg_bytes=seg, max_seg=rng.choice([8,32,128,512]), window=win)
            if not (st==1 and out==data and cons==len(z)): raise Exception("spec st %d len %d cons %d/%d %s"%(st,len(out),cons,len(z),stats))
    except Exception as e:
        fails+=1; print("INFLATE FAIL", it, n, zl, strat, repr(e)[:160]); open('/tmp/fuzz_fail_%d.z'%it,'wb').write(z)
print("iterations", it, "fails", fails, "seconds", round(time.time()-t0))
