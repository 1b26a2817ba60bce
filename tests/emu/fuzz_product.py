"""Randomized differential run of the WHOLE emulated product (tests/emu/libmz_strm_emu.so) through the vtbl against zlib:
read path (K5/K6, random read sizes, truncation), write path (random levels / write sizes / framings).
  MZ_CUDA_BATCH_KB=1024 MZ_CUDA_SPEC_SEG_KB=2 MZ_CUDA_READ_WINDOW_KB=1536 python tests/emu/fuzz_product.py <seed> <seconds>
TEST INFRASTRUCTURE ONLY."""
import sys, os, zlib, random, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuharness, datagen
p = cuharness.pkg()
lib = p.configure(C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libmz_strm_emu.so')))
tl = cuharness.TestLib()
rng = random.Random(int(sys.argv[1])); T=float(sys.argv[2])
def gen(n):
    k = rng.randrange(6)
    if k==0: return datagen.text_like(n, rng.randrange(1<<20))
    if k==1: return datagen.mixed(n, rng.randrange(1<<20))
    if k==2: return datagen.random_bytes(n, rng.randrange(1<<20))
    if k==3: return datagen.binary_records(n, rng.randrange(1<<20))
    if k==4: return bytes([rng.randrange(256)])*n
    parts=[]; left=n
    while left>0:
        m=min(left, rng.randrange(1,200000)); parts.append(gen(m)); left-=m
    return b"".join(parts)
t0=time.time(); it=0; fails=0
while time.time()-t0<T:
    it+=1
    n = rng.choice([0,1,100,65536,300000,1000000,2500000]) if rng.random()<0.4 else rng.randrange(0,3000000)
    data = gen(n)
    wbits = rng.choice([-15,15,31]); zl=rng.choice([0,1,6,9])
    co = zlib.compressobj(zl, zlib.DEFLATED, wbits, rng.choice([1,8,9]), rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_RLE, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY]))
    parts=[]; o=0
    while o<len(data):
        m=rng.randrange(1,400000); parts.append(co.compress(data[o:o+m])); o+=m
        if rng.random()<0.2: parts.append(co.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])))
    parts.append(co.flush()); z=b"".join(parts)
    rsize = rng.choice([1000,16384,65535,300000,1<<20])
    try:
        out, info = tl.decompress(lib.mz_stream_cuda_create, z + (bytes(rng.randrange(0,50)) if wbits==-15 else b""), len(data), window_bits=wbits, read_size=rsize)
        assert info["read"]==len(data) and out==data and info["total_in"]==len(z) and info["error"]==0 and info["close"]==0, info
        # write path round trip
        lvl = rng.choice([0,1,2,4,6,9]); ws = rng.choice([1000,16384,65535,500000])
        if n <= 1500000:
            comp, winfo = tl.compress(lib.mz_stream_cuda_create, data, level=lvl, window_bits=wbits, write_size=ws)
            assert winfo["close"]==0 and zlib.decompress(comp, wbits)==data and winfo["total_out"]==len(comp)
        # truncated / corrupted must not return wrong data silently
        if len(z) > 100:
            cut = rng.randrange(1, len(z))
            out, info = tl.decompress(lib.mz_stream_cuda_create, z[:cut], len(data), window_bits=wbits, read_size=rsize)
            got = out or b""
            assert data.startswith(got), "truncated produced wrong bytes"
            assert info["error"] != 0 or (len(got)==len(data)), ("truncated", cut, len(z), info)
    except Exception as e:
        fails+=1; print("FAIL", it, n, wbits, zl, rsize, repr(e)[:300]); open('/tmp/fuzz_host_fail_%d.z'%it,'wb').write(z)
print("iterations", it, "fails", fails, "seconds", round(time.time()-t0))
