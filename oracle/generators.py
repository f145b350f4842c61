"""TEST INFRASTRUCTURE - CPU restatement of generators 2..9 of the reference's tools/generators.c (generator 1, the
LCG, is in oracle/sts_oracle.c).  Only tests/, smoke() and bench.py's baseline legs may import this.

Each function returns the FILE BYTES the tool writes for `streams` consecutive streams of n bits (n = 2^20 in the
tool, generators.c:141; the restatement takes any n): write_sequence (:1603-1660) packs sequence bit 8j + c into
bit c of byte j, LSB first.  copyBitsToEpsilon (:704-760) cuts the generator's blocks into streams: a stream takes
whole blocks until it has n bits and DROPS the rest of its last block; the state carries over to the next stream.

The tool's arithmetic is on big-endian byte strings (Mult :1899, ModMult :1936, ModExp :2003, DivMod :2043).  Where
it is a correct big-number routine the restatement uses Python integers; where it is not, the byte-level behaviour
is restated, because THAT is what the tool's output is:

  Q1  smult (:1968-1979) adds its carry twice and skips the top byte of the multiplicand, and quadRes2 (:996-1002)
      uses the 64-byte result inside a 65-byte number, so "2x + 3" is not what is computed.
  Q2  bbs (:1288) squares IN PLACE: Mult's accumulator A aliases both factors and is not cleared, so each row of the
      schoolbook product reads bytes earlier rows have overwritten.  The map is deterministic but it is not
      x -> x^2 mod n, and there is no way to jump ahead in it.
  Q3  bbs prints its "BITSREAD" progress lines to stdout (:1300), into the data; the restatement returns data only.
  Q4  ModExp (:2012) sets only the LAST byte of its accumulator to 1.  modExp clears the accumulator before every
      call (:1224); micali_schnorr does not (:1380): block t starts from the previous block's result with its last
      byte replaced by 1, and the FIRST block starts from uninitialised stack memory - the tool's generator 8 output
      differs from run to run.  `ms_initial` is that memory (128 bytes); the engine defines it as zero.
  Q5  ahtopb (:2447) reads two characters per byte: the 47-digit seed of micali_schnorr (:1375) ends in the
      string's NUL, which makes the last byte 0x30.
  Q6  SHA1 (:1484, :1579) reverses bytes only `#ifdef LITTLE_ENDIAN`, which the tool's own Makefile flags
      (-std=c99 -D_ISOC99_SOURCE, tools/Makefile:59) leave undefined on glibc: on x86 the message words and the
      output words are little-endian, so this is not SHA-1.  Restated as built by the reference's Makefile.
  Q7  exclusiveOR (:1141-1154) seeds its ring with the CHARACTERS '0'/'1' and stores new bits as the numbers 0/1;
      while the two kinds are compared the "xor" is always 1: bits 128..253 of the sequence are all ones.
  Q8  cubicRes (:1075-1078) emits the TOP 512 bits of the 1536-bit cube and keeps the low 512 as the state.
"""
import numpy as np

_P512 = int("987b6a6bf2c56a97291c445409920032499f9ee7ad128301b5d0254aa1a9633f"
            "dbd378d40149f1e23a13849f3d45992f5c4c6b7104099bc301f6005f9d8115e1", 16)      # :942, :1213
_G1 = int("3844506a9456c564b8b8538e0cc15aff46c95e69600f084f0657c2401b3c2447"
          "34b62ea9bb95be4923b9b7e84eeaf1a224894ef0328d44bc3eb3e983644da3f5", 16)        # :945, :1216
_G2 = int("7844506a9456c564b8b8538e0cc15aff46c95e69600f084f0657c2401b3c2447"
          "34b62ea9bb95be4923b9b7e84eeaf1a224894ef0328d44bc3eb3e983644da3f5", 16)        # :1010, :1070
_BBS_P = int("E65097BAEC92E70478CAF4ED0ED94E1C94B154466BFB9EC9BE37B2B0FF8526C2"
             "22B76E0E915017535AE8B9207250257D0A0C87C0DACEF78E17D1EF9DC44FD91F", 16)     # :1274, :1364
_BBS_Q = int("E029AEFCF8EA2C29D99CB53DD5FA9BC1D0176F5DF8D9110FD16EE21F32E37BA8"
             "6FF42F00531AD5B8A43073182CC2E15F5C86E8DA059E346777C9A985F7D8A867", 16)     # :1277, :1367
_BBS_S = int("10d6333cfac8e30e808d2192f7c0439480da79db9bbca1667d73be9a677ed313"
             "11f3b830937763837cb7b1b1dc75f14eea417f84d9625628750de99e7ef1e976", 16)     # :1283
_MODEXP_Y = int("7AB36982CE1ADF832019CDFEB2393CABDF0214EC", 16)                          # :1212
_MS_X = (int("237c5f791c2cfe47bfb16d2d54a0d60665b20904ec822a", 16) << 8) | 0x30          # :1375 through Q5
_SHA_XKEY = int("ec822a619d6ed5d9492218a7a4c5b15d57c61601", 16)                          # :1459
_XOR_SEED = ("0001011011011001000101111001001010011011101101000100000010101111"
             "111010100100001010110110000000000100110000101110011111111100111")         # :1134-1136

NAMES = {2: "QCG1", 3: "QCG2", 4: "CCG", 5: "XOR", 6: "MODEXP", 7: "BBS", 8: "MS", 9: "G-SHA1"}   # :557-568


def _pack(bits):
    return np.packbits(np.asarray(bits, dtype=np.uint8), bitorder="little").tobytes()


def _bits(x, nbytes, nbits):
    """the first nbits of the nbytes big-endian bytes of x, MSB of each byte first (copyBitsToEpsilon :730-752)"""
    return np.unpackbits(np.frombuffer(x.to_bytes(nbytes, "big"), dtype=np.uint8))[:nbits]


def _cut(block, n, streams):
    out = []
    for _ in range(streams):
        cur, got = [], 0
        while got < n:
            b = block()
            take = min(len(b), n - got)
            cur.append(b[:take])
            got += take
        out.append(_pack(np.concatenate(cur)))
    return out


def quadres1(n, streams):
    """generators.c:915-975: g <- g^2 mod p, the 512 bits of g are the block"""
    st = [_G1]

    def block():
        st[0] = st[0] * st[0] % _P512
        return _bits(st[0], 64, 512)
    return _cut(block, n, streams)


def quadres2_step(g):
    """generators.c:996-1003 with Q1"""
    c = g.to_bytes(64, "big")
    a = bytearray(65)
    res = 0
    for i in range(63, 0, -1):                          # smult :1975-1979
        res = (a[i] + 2 * c[i] + (res >> 8)) & 0xFFFF
        a[i] = res & 0xFF
        a[i - 1] = res >> 8
    t1 = int.from_bytes(a, "big") + 3                    # add(t1, 65, Three, 1): no carry out of 65 bytes here
    return (t1 * g + 1) % (1 << 512)


def quadres2(n, streams):
    st = [_G2]

    def block():
        st[0] = quadres2_step(st[0])
        return _bits(st[0], 64, 512)
    return _cut(block, n, streams)


def cubicres(n, streams):
    """generators.c:1049-1108 with Q8"""
    st = [_G2]

    def block():
        c = st[0] ** 3
        st[0] = c % (1 << 512)
        return _bits(c >> 1024, 64, 512)
    return _cut(block, n, streams)


def xor_bits(total):
    """generators.c:1111-1186 with Q7: the first `total` bits of the one sequence all streams are cut from"""
    b = np.zeros(max(total, 254), dtype=np.uint8)
    b[:127] = [int(ch) for ch in _XOR_SEED]
    b[127] = b[126] ^ b[0]                              # the one comparison between two characters
    b[128:254] = 1
    for i in range(254, total):
        b[i] = b[i - 1] ^ b[i - 127]
    return b[:total]


def exclusive_or(n, streams):
    b = xor_bits(n * streams)
    return [_pack(b[k * n:(k + 1) * n]) for k in range(streams)]


def modexp(n, streams):
    """generators.c:1192-1253: x = g^y mod p is the block, y <- the low 160 bits of x"""
    y = [_MODEXP_Y]

    def block():
        x = pow(_G1, y[0], _P512)
        y[0] = x % (1 << 160)
        return _bits(x, 64, 512)
    return _cut(block, n, streams)


def bbs_step(x, low, nn):
    """one pass of the loop body generators.c:1288-1289 with Q2.  `x` is bytes 0..127 of the tool's buffer (the value
    to be squared), `low` bytes 128..255 (what the previous pass left there); returns the new value"""
    a = bytearray(x.to_bytes(128, "big") + low.to_bytes(128, "big"))
    for i in range(127, -1, -1):                        # Mult :1908-1917, one row per byte of B, C read as of now
        c = int.from_bytes(a[0:128], "big")
        t = int.from_bytes(a[i + 1:i + 129], "big") + a[i] * c
        a[i:i + 129] = t.to_bytes(129, "big")           # A[--k] = carry: assigned, not added
    return int.from_bytes(a, "big") % nn


def bbs(n, streams):
    nn = _BBS_P * _BBS_Q
    x = _BBS_S * _BBS_S % nn                            # :1286: not in place, a true square
    low = 0
    out = []
    for _ in range(streams):
        b = np.zeros(n, dtype=np.uint8)
        for i in range(n):
            x = bbs_step(x, low, nn)
            low = x                                     # the remainder stays in bytes 128..255 (Mod :2194-2196)
            b[i] = x & 1
        out.append(_pack(b))
    return out


def micali_schnorr(n, streams, ms_initial=0):
    """generators.c:1346-1425 with Q4 and Q5; e = 11, k = 837 output bits and r = 187 state bits per block"""
    nn = _BBS_P * _BBS_Q
    st = [_MS_X, ms_initial]

    def block():
        a = (st[1] >> 8 << 8) | 1                       # ModExp :2012
        for bit in (1, 0, 1, 1):                        # e = 0x0b from its top set bit (:2014-2037)
            a = a * a % nn
            if bit:
                a = a * st[0] % nn
        st[1] = a
        tail = (a % (1 << 840)) << 3 & ((1 << 840) - 1)  # Y + 23, 105 bytes, shifted left 3 (:1382-1385)
        st[0] = (a >> (1024 - 192)) >> 5                # the top 24 bytes shifted right 5 (:1388-1392)
        return _bits(tail, 105, 837)
    return _cut(block, n, streams)


def _rol(v, s):
    return ((v << s) | (v >> (32 - s))) & 0xFFFFFFFF


def _sha_g(c):
    """generators.c:1474-1582 with Q6: the compression function over Xkey || 0^352 with little-endian words"""
    m = c.to_bytes(20, "big") + bytes(44)
    w = [int.from_bytes(m[4 * i:4 * i + 4], "little") for i in range(16)]
    for t in range(16, 80):
        w.append(_rol(w[t - 3] ^ w[t - 8] ^ w[t - 14] ^ w[t - 16], 1))
    h = [0x67452301, 0xEFCDAB89, 0x98BADCFE, 0x10325476, 0xC3D2E1F0]
    a, b, cc, d, e = h
    for t in range(80):
        if t < 20:
            f, k = (b & cc) | (~b & d), 0x5A827999
        elif t < 40:
            f, k = b ^ cc ^ d, 0x6ED9EBA1
        elif t < 60:
            f, k = (b & cc) | (b & d) | (cc & d), 0x8F1BBCDC
        else:
            f, k = b ^ cc ^ d, 0xCA62C1D6
        tmp = (_rol(a, 5) + (f & 0xFFFFFFFF) + e + k + w[t]) & 0xFFFFFFFF
        e, d, cc, b, a = d, cc, _rol(b, 30), a, tmp
    r = [(x + y) & 0xFFFFFFFF for x, y in zip(h, [a, b, cc, d, e])]
    return int.from_bytes(b"".join(x.to_bytes(4, "little") for x in r), "big")


def sha1(n, streams):
    """generators.c:1428-1600: block = G(Xkey) (160 bits), Xkey <- Xkey + G + 1 mod 2^160 (:1585-1586)"""
    xk = [_SHA_XKEY]

    def block():
        g = _sha_g(xk[0])
        xk[0] = (xk[0] + g + 1) % (1 << 160)
        return _bits(g, 20, 160)
    return _cut(block, n, streams)


def generate(gen, n, streams, ms_initial=0):
    """the file bytes of `generators <gen> <streams>` (for n = 2^20), one bytes object per stream"""
    f = {2: quadres1, 3: quadres2, 4: cubicres, 5: exclusive_or, 6: modexp, 7: bbs, 9: sha1}
    if gen == 8:
        return micali_schnorr(n, streams, ms_initial)
    return f[gen](n, streams)
