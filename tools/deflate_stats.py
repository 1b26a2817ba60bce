#!/usr/bin/env python3
"""Walk a raw DEFLATE stream block by block and report, per dynamic block, what its symbols cost under the code the
encoder chose against what an optimal (package-merge-free: plain Huffman, then verified <= 15 bits) code would cost.
Design tool for the code builder of the deflate kernel; pure Python, a few MB at most."""
import heapq
import sys

LBASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEXT = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
DEXT = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]
ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


class Bits:
    def __init__(self, data):
        self.v = int.from_bytes(data, "little")
        self.pos = 0

    def get(self, n):
        r = (self.v >> self.pos) & ((1 << n) - 1)
        self.pos += n
        return r


def mktable(lens):
    """canonical code -> dict (len, code MSB-first) -> symbol"""
    cnt = [0] * 16
    for l in lens:
        cnt[l] += 1
    cnt[0] = 0
    nxt, code = [0] * 16, 0
    for b in range(1, 16):
        code = (code + cnt[b - 1]) << 1
        nxt[b] = code
    t = {}
    for s, l in enumerate(lens):
        if l:
            t[(l, nxt[l])] = s
            nxt[l] += 1
    return t


def decode(bits, table):
    code, n = 0, 0
    while True:
        code = (code << 1) | bits.get(1)
        n += 1
        if (n, code) in table:
            return table[(n, code)]
        if n > 15:
            raise ValueError("bad code")


def huff_cost(freq, maxbits=15):
    items = [(f, i) for i, f in enumerate(freq) if f]
    if len(items) < 2:
        return sum(freq)
    h = [(f, i, None) for f, i in items]
    heapq.heapify(h)
    uid = len(freq)
    depth = {}
    nodes = {}
    while len(h) > 1:
        a = heapq.heappop(h)
        b = heapq.heappop(h)
        uid += 1
        nodes[uid] = (a, b)
        heapq.heappush(h, (a[0] + b[0], uid, True))
    cost = 0
    stack = [(h[0], 0)]
    mx = 0
    while stack:
        (f, i, inner), d = stack.pop()
        if inner:
            a, b = nodes[i]
            stack.append((a, d + 1))
            stack.append((b, d + 1))
        else:
            cost += f * d
            mx = max(mx, d)
    return cost  # (blocks of 32 KiB rarely exceed 15 bits; the figure is a lower bound when they do)


def main():
    data = open(sys.argv[1], "rb").read()
    bits = Bits(data)
    tot_actual = tot_opt = tot_hdr = tot_extra = nblk = 0
    while True:
        final, typ = bits.get(1), bits.get(2)
        if typ == 0:
            bits.pos = (bits.pos + 7) & ~7
            n = bits.get(16)
            bits.get(16)
            bits.pos += 8 * n
        elif typ == 2:
            start = bits.pos - 3
            hlit, hdist, hclen = bits.get(5) + 257, bits.get(5) + 1, bits.get(4) + 4
            cl = [0] * 19
            for i in range(hclen):
                cl[ORDER[i]] = bits.get(3)
            ct = mktable(cl)
            lens = []
            while len(lens) < hlit + hdist:
                s = decode(bits, ct)
                if s < 16:
                    lens.append(s)
                elif s == 16:
                    lens += [lens[-1]] * (3 + bits.get(2))
                elif s == 17:
                    lens += [0] * (3 + bits.get(3))
                else:
                    lens += [0] * (11 + bits.get(7))
            hdr = bits.pos - start
            ll, dl = lens[:hlit], lens[hlit:]
            lt, dt = mktable(ll), mktable(dl)
            fl, fd = [0] * 286, [0] * 30
            extra = 0
            while True:
                s = decode(bits, lt)
                fl[s] += 1
                if s == 256:
                    break
                if s > 256:
                    e = LEXT[s - 257]
                    bits.get(e)
                    d = decode(bits, dt)
                    fd[d] += 1
                    bits.get(DEXT[d])
                    extra += e + DEXT[d]
            actual = sum(f * ll[i] for i, f in enumerate(fl) if f) + sum(f * dl[i] for i, f in enumerate(fd) if f)
            opt = huff_cost(fl) + huff_cost(fd)
            tot_actual += actual
            tot_opt += opt
            tot_hdr += hdr
            tot_extra += extra
            nblk += 1
        else:
            raise ValueError("fixed block")
        if final:
            break
    n = max(1, int(sys.argv[2])) if len(sys.argv) > 2 else 1
    print("blocks %d  code bits %d  optimal %d  (+%.3f %%)  header %d  extra %d  | of input: codes %.4f opt %.4f hdr %.4f extra %.4f" %
          (nblk, tot_actual, tot_opt, 100.0 * (tot_actual - tot_opt) / tot_opt, tot_hdr, tot_extra,
           tot_actual / 8 / n, tot_opt / 8 / n, tot_hdr / 8 / n, tot_extra / 8 / n))


if __name__ == "__main__":
    main()
