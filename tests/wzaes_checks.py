"""Shared checks for K8 (WinZip AES arithmetic, scope row f4): known answers from independent implementations -- hashlib's
PBKDF2 / HMAC, the `cryptography` package's AES (ECB of the counter blocks = the CTR key stream, little-endian counter from 1 as
mz_strm_wzaes.c:147-171 builds it) -- for every strength, odd lengths, unaligned offsets, a password longer than one SHA-1 block.
`call(name, *args)` invokes the product entry point; `buf(bytes)` gives (pointer, keepalive) for device-visible memory and
`back(ptr, n)` reads it back."""
import hashlib
import hmac
import os


def ctr_keystream_xor(key, data):
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    enc = Cipher(algorithms.AES(key), modes.ECB()).encryptor()
    ks = enc.update(b"".join((i + 1).to_bytes(8, "little") + bytes(8) for i in range((len(data) + 15) // 16)))
    return bytes(a ^ b for a, b in zip(data, ks))


def run(call, buf, back, lens=(0, 1, 15, 16, 17, 4097, 100001)):
    import ctypes as C
    rng = os.urandom
    for strength in (1, 2, 3):
        klen, slen = 8 * strength + 8, 4 * strength + 4
        n = len(lens)
        pw = b"correct horse battery" if strength < 3 else bytes(range(33, 33 + 100))  # > 64 bytes: the key is hashed first
        salts = rng(16 * n)
        p_pw, k1 = buf(pw)
        p_salt, k2 = buf(salts)
        p_keys, k3 = buf(bytes(80 * n))
        assert call("mz_cuda_wzaes_derive", p_pw, len(pw), p_salt, n, strength, p_keys, None) == 0
        keys = back(p_keys, 80 * n)
        for e in range(n):
            dk = hashlib.pbkdf2_hmac("sha1", pw, salts[16 * e:16 * e + slen], 1000, 2 * klen + 2)
            r = keys[80 * e:80 * e + 80]
            assert r[:klen] == dk[:klen] and r[32:32 + klen] == dk[klen:2 * klen] and r[64:66] == dk[2 * klen:], (strength, e)
        offs, pos = [], 3
        for ln in lens:
            offs.append(pos)
            pos += ln + 5
        plain = rng(pos)
        p_data, k4 = buf(plain)
        p_off, k5 = buf(b"".join(o.to_bytes(8, "little") for o in offs))
        p_len, k6 = buf(b"".join(ln.to_bytes(8, "little") for ln in lens))
        p_mac, k7 = buf(bytes(20 * n))
        assert call("mz_cuda_wzaes_ctr", p_data, p_off, p_len, n, max(lens), p_keys, strength, None) == 0
        assert call("mz_cuda_wzaes_hmac", p_data, p_off, p_len, n, p_keys, strength, p_mac, None) == 0
        data = back(p_data, pos)
        mac = back(p_mac, 20 * n)
        for e in range(n):
            r = keys[80 * e:80 * e + 80]
            ct = data[offs[e]:offs[e] + lens[e]]
            assert ct == ctr_keystream_xor(r[:klen], plain[offs[e]:offs[e] + lens[e]]), (strength, e, "ctr")
            assert mac[20 * e:20 * e + 20] == hmac.new(r[32:32 + klen], ct, hashlib.sha1).digest(), (strength, e, "hmac")
            end = offs[e] + lens[e]
            assert data[end:end + 5] == plain[end:end + 5]  # bytes between the entries are not touched
        # applying the key stream again decrypts
        assert call("mz_cuda_wzaes_ctr", p_data, p_off, p_len, n, max(lens), p_keys, strength, None) == 0
        assert back(p_data, pos) == plain
    # parameter checks
    assert call("mz_cuda_wzaes_derive", None, 0, None, 1, 0, None, None) != 0
    assert call("mz_cuda_wzaes_ctr", None, None, None, 1, 16, None, 4, None) != 0


def aes_roundtrip(exe_e, refe, tmp_path, run_cmd, n=60, esz=30_000):
    """archives both ways against the reference built with WinZip AES (oracle/_ref/minizip_refe):
    ours (native writer + K8) -> the reference CLI extracts with the password and the reference's own loop decrypts every entry;
    the reference's AES archive -> the batch extractor (K8 + K5); wrong password / damaged ciphertext are refused with the
    reference's codes"""
    import json
    import zipfile
    (tmp_path / "dumpa").mkdir()
    st = json.loads(run_cmd([exe_e, "a.zip", str(n), str(esz), "6", "native_aes", "dumpa", "7"], tmp_path).stdout.decode().strip().splitlines()[-1])
    assert st["err"] == 0 and st["entries"] == n
    with zipfile.ZipFile(tmp_path / "a.zip") as z:
        for i in z.infolist():
            assert i.compress_type == 99 and i.flag_bits & 0x809 == 0x801 and i.extract_version == 51
            assert i.extra[:11] == bytes([0x01, 0x99, 7, 0, 1, 0, 0x41, 0x45, 3, 8, 0]), i.extra.hex()
    run_cmd([refe, "-x", "-o", "-p", "secret", "-d", "outa", "a.zip"], tmp_path)
    for k in range(0, n, 7):
        assert (tmp_path / "outa" / "e" / ("%06d" % k)).read_bytes() == (tmp_path / "dumpa" / ("%06d" % k)).read_bytes()
    got = json.loads(run_cmd([exe_e, "a.zip", str(n), str(esz), "0", "extract_aes"], tmp_path).stdout.decode().strip().splitlines()[-1])
    want = json.loads(run_cmd([exe_e, "a.zip", str(n), str(esz), "0", "extract_ref_aes"], tmp_path).stdout.decode().strip().splitlines()[-1])
    assert got["err"] == 0 and want["err"] == 0 and got["entries"] == want["entries"] == n and got["bytes_out"] == want["bytes_out"] and got["mismatches"] == 0
    bad = json.loads(run_cmd(["env", "ZIPBATCH_PASSWORD=wrong", exe_e, "a.zip", str(n), str(esz), "0", "extract_aes"], tmp_path, ok=False).stdout.decode().strip().splitlines()[-1])
    assert bad["err"] == -108  # MZ_PASSWORD_ERROR
    # an archive the REFERENCE encrypted (OpenSSL), incl. an empty entry and AES-128 is not an option of its CLI: strength 256
    src = tmp_path / "srca"
    src.mkdir()
    files = {"t1.txt": (b"the quick brown fox " * 9000), "t2.bin": os.urandom(3000), "empty.bin": b""}
    for k, v in files.items():
        (src / k).write_bytes(v)
    run_cmd([refe, "-s", "-p", "secret", "-o", "../r_a.zip"] + sorted(files), src)
    got = json.loads(run_cmd([exe_e, "r_a.zip", "3", str(1 << 20), "0", "extract_aes"], tmp_path).stdout.decode().strip().splitlines()[-1])
    assert got["err"] == 0 and got["entries"] == 3 and got["bytes_out"] == sum(len(v) for v in files.values())
    # one flipped ciphertext bit: the HMAC no longer matches (MZ_CRC_ERROR, as mz_stream_wzaes_close reports it)
    raw = bytearray((tmp_path / "a.zip").read_bytes())
    with zipfile.ZipFile(tmp_path / "a.zip") as z:
        victim = z.infolist()[n // 2]
    raw[victim.header_offset + 30 + len(victim.filename) + len(victim.extra) + 18 + 100] ^= 1
    (tmp_path / "a_bad.zip").write_bytes(bytes(raw))
    bad = json.loads(run_cmd([exe_e, "a_bad.zip", str(n), str(esz), "0", "extract_aes"], tmp_path, ok=False).stdout.decode().strip().splitlines()[-1])
    assert bad["err"] == -105  # MZ_CRC_ERROR
