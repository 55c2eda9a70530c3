"""Generates tests/golden/c2df_fuzz_golden.npz: ~2000 damaged / unusual ``.c2df`` files together with what THE
REFERENCE'S OWN CODE does with each of them (src/filemaker.py ``unpack_c2df`` + src/search.py
``decode_clip_from_c2df``, run unmodified in the build container through the stubs of make_golden.py).
Re-run:  python tests/golden/make_fuzz_golden.py

build.py:80-88 skips a file when that call raises and keeps it otherwise, so the contract of the batched walker
(csrc/c2df_walk.cpp, sgic_c2df_parse) is: status 0 with the same u8 codes exactly for the files the reference keeps,
a non-zero status for exactly the files it skips.  The reference loads EVERY entry of a file eagerly
(filemaker.py:102-135: JSON entries are parsed, strings decoded, arrays rebuilt) and parses the header JSON, so damage
anywhere in a file — not only in ``clip_stream`` / ``clip_meta`` — decides whether the file is skipped.

Mutations (fixed seed): byte flips anywhere / inside the JSON regions, truncations, trailing junk, hand-written JSON
texts for ``clip_meta`` / the header / other entries (every token class of CPython's json scanner, UTF-8 edge cases,
escapes, duplicate keys, ``dim`` spellings), keys and strings that are not UTF-8, unknown type codes, damaged array
entries, length fields that lie, duplicate entries, entries cut off by the end of the file.
"""
import struct
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as mg  # noqa: E402  (stubs + the reference modules: mg.ref_fm, mg.ref_search)

OUT = HERE / "c2df_fuzz_golden.npz"
D = 64


def entry(key, t, payload, *, raw_key=None):
    k = key.encode("utf-8") if raw_key is None else raw_key
    head = struct.pack("<H", len(k)) + k + struct.pack("<B", t)
    if t in (2, 3, 7, 6):
        return head + payload
    return head + struct.pack("<I", len(payload)) + payload


def lp(b):  # BYTES / STR / JSON payloads carry their own length again
    return struct.pack("<I", len(b)) + b


def np_payload(dt=b"<i4", shape=(2,), data=None, ndim=None, data_len=None):
    data = np.arange(int(np.prod(shape)), dtype="<i4").tobytes() if data is None else data
    out = struct.pack("<B", len(dt)) + dt + struct.pack("<B", len(shape) if ndim is None else ndim)
    for s in shape:
        out += struct.pack("<I", s)
    return out + struct.pack("<I", len(data) if data_len is None else data_len) + data


def raw_file(entries, header=b'{"version": 2}', n_items=None, magic=b"C2DF", hlen=None):
    body = b"".join(entries)
    return (magic + struct.pack("<H", 1) + struct.pack("<I", len(header) if hlen is None else hlen) + header +
            struct.pack("<I", len(entries) if n_items is None else n_items) + body)


def main():
    rng = np.random.default_rng(20261019)
    v = rng.standard_normal(D).astype(np.float32)
    v /= np.linalg.norm(v)
    codes = mg.quantize(v)
    frame = mg.ZstdCompressor(level=19).compress(codes.tobytes())
    meta_ok = b'{"model_id": "ViT-B-32:laion2b_s34b_b79k", "dim": 64, "quant": "u8_symmetric_-1_1", "codec": "zstd", "zstd_level": 19}'
    E_STREAM = entry("clip_stream", 0, lp(frame))
    E_META = entry("clip_meta", 4, lp(meta_ok))

    def with_meta(text, **kw):
        return raw_file([E_STREAM, entry("clip_meta", 4, lp(text))], **kw)

    cases = []  # (description, bytes)

    def add(desc, b):
        cases.append((desc, bytes(b)))

    # ---- seeds made by the reference's packer --------------------------------------------------------------
    seeds = [mg.blob_for(v, rng, big=False), mg.blob_for(v, rng, big=True),
             mg.blob_for(v, rng, big=True, extra={"b_str": "héllo ✓", "c_int": -5, "d_float": 2.5, "h_none": None,
                                                    "i_bool": True, "e_json": {"k": [1, 2.5, "x", None, True]},
                                                    "g_np": np.arange(12, dtype=np.float32).reshape(3, 4)})]
    for i, s in enumerate(seeds):
        add(f"seed{i}", s)
    add("minimal raw", raw_file([E_STREAM, E_META]))

    # ---- clip_meta JSON texts: every token class of the scanner ---------------------------------------------
    metas = [
        b'{"dim": 64}', b' \t\n\r{ "dim" : 64 } \n', b'{"dim":64,}', b'{"dim":64,"x":}', b'{"dim":64 "x":1}', b'{,"dim":64}',
        b'{"dim":64}x', b'{"dim":64}{}', b'{"dim":64} ,', b'', b' ', b'{', b'}', b'{"dim":64', b'{"dim"}', b'{"dim":}',
        b"{'dim':64}", b'{dim:64}', b'{"dim":64,"a":tru}', b'{"dim":64,"a":true}', b'{"dim":64,"a":True}', b'{"dim":64,"a":nul}',
        b'{"dim":64,"a":null}', b'{"dim":64,"a":falsey}', b'{"dim":64,"a":NaN}', b'{"dim":64,"a":Infinity}',
        b'{"dim":64,"a":-Infinity}', b'{"dim":64,"a":-Inf}', b'{"dim":64,"a":+1}', b'{"dim":64,"a":01}', b'{"dim":64,"a":-0}',
        b'{"dim":64,"a":-}', b'{"dim":64,"a":1.}', b'{"dim":64,"a":.5}', b'{"dim":64,"a":1.5e}', b'{"dim":64,"a":1.5e+}',
        b'{"dim":64,"a":1.5e+3}', b'{"dim":64,"a":1E-2}', b'{"dim":64,"a":1e5x}', b'{"dim":64,"a":0x10}', b'{"dim":64,"a":1_0}',
        b'{"dim":64,"a":[1,2,]}', b'{"dim":64,"a":[,1]}', b'{"dim":64,"a":[1 2]}', b'{"dim":64,"a":[]}', b'{"dim":64,"a":[ ]}',
        b'{"dim":64,"a":{}}', b'{"dim":64,"a":{ }}', b'{"dim":64,"a":[[[[{"b":[{}]}]]]]}', b'{"dim":64,"a":[1,[2,{"dim":9}]]}',
        b'{"dim":64,"a":"\\n\\t\\"\\\\\\/\\b\\f\\r"}', b'{"dim":64,"a":"\\x41"}', b'{"dim":64,"a":"\\u0041"}', b'{"dim":64,"a":"\\u00"}',
        b'{"dim":64,"a":"\\u00zz"}', b'{"dim":64,"a":"\\ud800"}', b'{"dim":64,"a":"\\ud83d\\ude00"}', b'{"dim":64,"a":"\\"}',
        b'{"dim":64,"a":"tab\there"}', b'{"dim":64,"a":"nl\nhere"}', b'{"dim":64,"a":"\x7f"}', b'{"dim":64,"a":"\x1f"}',
        b'{"dim":64,"a":"unterminated}', b'{"dim":64,"a":"\xc3\xa9"}', b'{"dim":64,"a":"\xc3"}', b'{"dim":64,"a":"\xc0\xaf"}',
        b'{"dim":64,"a":"\xe0\x80\xaf"}', b'{"dim":64,"a":"\xed\xa0\x80"}', b'{"dim":64,"a":"\xf4\x90\x80\x80"}',
        b'{"dim":64,"a":"\xf0\x9f\x98\x80"}', b'{"dim":64,"a":"\xf8\x88\x80\x80\x80"}', b'{"dim":64,"a":"\x80"}',
        b'{"dim":64,"a":"\xe2\x82"}', b'{"dim":64,"a":\xc3\xa9}', b'\xef\xbb\xbf{"dim":64}', b'{"dim":64}\xc2\xa0', b'\x0c{"dim":64}',
        b'{"dim":64}\x00', b'{"d\\u0069m":64}', b'{"dim":1,"dim":64}', b'{"dim":64,"dim":1}', b'{"dim":"abc","dim":64}',
        b'{"dim":64,"dim":"abc"}', b'{"dim":[64],"dim":64}', b'{"Dim":64}', b'{"dim ":64}', b'{"a":{"dim":64}}', b'[{"dim":64}]',
        b'[]', b'[64]', b'null', b'false', b'true', b'0', b'64', b'""', b'"dim"', b'{}', b'{ }', b'NaN',
        b'{"dim":64.0}', b'{"dim":64.9}', b'{"dim":63.999999999999999}', b'{"dim":6.4e1}', b'{"dim":6400e-2}', b'{"dim":0.64E2}',
        b'{"dim":-64}', b'{"dim":0}', b'{"dim":-0}', b'{"dim":064}', b'{"dim":64e0}', b'{"dim":1e400}', b'{"dim":-1e400}',
        b'{"dim":NaN}', b'{"dim":Infinity}', b'{"dim":true}', b'{"dim":false}', b'{"dim":null}', b'{"dim":[64]}', b'{"dim":{"dim":64}}',
        b'{"dim":"64"}', b'{"dim":" 64 "}', b'{"dim":"\\t64\\n"}', b'{"dim":"+64"}', b'{"dim":"-64"}', b'{"dim":"6_4"}', b'{"dim":"_64"}',
        b'{"dim":"64_"}', b'{"dim":"6__4"}', b'{"dim":"064"}', b'{"dim":"0064"}', b'{"dim":"64.0"}', b'{"dim":"6e1"}', b'{"dim":"0x40"}',
        b'{"dim":""}', b'{"dim":" "}', b'{"dim":"+"}', b'{"dim":"6 4"}', b'{"dim":"\\u0036\\u0034"}', b'{"dim":"\\u001c64\\u001f"}',
        b'{"dim":"64\\u0000"}', b'{"dim":"' + b"0" * 40 + b'64"}', b'{"dim":' + b"9" * 30 + b'}', b'{"dim":64' + b" " * 300 + b'}',
        b'{"dim":18446744073709551680}', b'{"dim":4294967360}', b'{"dim":64,"big":' + b"1" * 4400 + b'}',
        b'{"dim":64,"big":' + b"1" * 4300 + b'}', b'{"dim":64,"a":' + b"[" * 200 + b"]" * 200 + b'}',
        b'{"dim":64,"a":' + b"[" * 200 + b"]" * 199 + b'}', b'{"dim":64,"a":1e999999}', b'{"dim":64,"a":-1.5E-999}',
    ]
    for t in metas:
        add("meta " + t[:60].decode("latin-1"), with_meta(t))
    # clip_meta of another type (falsy -> {} -> dim 0; truthy -> no .get)
    for t, p in ((6, b""), (7, b"\x00"), (7, b"\x01"), (2, struct.pack("<q", 0)), (2, struct.pack("<q", 64)),
                 (3, struct.pack("<d", 0.0)), (1, lp(b"")), (1, lp(b'{"dim":64}')), (0, lp(b"")), (0, lp(b'{"dim":64}')),
                 (5, np_payload())):
        add(f"meta as type {t}", raw_file([E_STREAM, entry("clip_meta", t, p)]))

    # ---- header texts -------------------------------------------------------------------------------------
    for h in (b"", b"{}", b"[]", b"null", b"7", b'"x"', b"{", b'{"a":}', b'{"a":1}x', b" {} ", b'{"note":"\xc3\xa9"}',
              b'{"note":"\xff"}', b"\xef\xbb\xbf{}", b'{"a":tru}', b'{"a":1,}', b"nul", b'{"a":"\x01"}', b"NaN", b"-Infinity", b"-"):
        add("header " + h[:40].decode("latin-1"), raw_file([E_STREAM, E_META], header=h))
    add("header length beyond the file", raw_file([E_STREAM, E_META], hlen=1 << 20))
    add("header length cuts the header", raw_file([E_STREAM, E_META], header=b'{"version": 2}', hlen=5))

    # ---- other entries: the reference loads them all --------------------------------------------------------
    others = {
        "str ok": entry("s", 1, lp("héllo".encode())), "str bad utf8": entry("s", 1, lp(b"\xff\xfe")),
        "str overlong": entry("s", 1, lp(b"\xc0\x80")), "str surrogate": entry("s", 1, lp(b"\xed\xb0\x80")),
        "str truncated seq": entry("s", 1, lp(b"ab\xe2\x82")), "str inner length lies": entry("s", 1, struct.pack("<I", 999) + b"abc"),
        "str no inner length": entry("s", 1, b"ab"), "bytes no inner length": entry("b", 0, b"abc"),
        "bytes inner length lies": entry("b", 0, struct.pack("<I", 999) + b"abc"), "bytes empty payload": entry("b", 0, b""),
        "json ok": entry("j", 4, lp(b"[1, 2, 3]")), "json bad": entry("j", 4, lp(b"[1, 2,")), "json not utf8": entry("j", 4, lp(b'"\xff"')),
        "json empty": entry("j", 4, lp(b"")), "json no inner length": entry("j", 4, b"[]"), "json inner length cuts": entry("j", 4, struct.pack("<I", 2) + b"[1]"),
        "json inner length lies": entry("j", 4, struct.pack("<I", 99) + b"[1]"),
        "int": entry("i", 2, struct.pack("<q", -1)), "float nan": entry("f", 3, struct.pack("<d", float("nan"))),
        "bool 2": entry("t", 7, b"\x02"), "none": entry("n", 6, b""),
        "type 8": entry("u", 8, lp(b"abc")), "type 255": entry("u", 255, lp(b"abc")), "type 9 empty": entry("u", 9, b""),
        "key bad utf8": entry("", 0, lp(b"x"), raw_key=b"\xff\xfe"), "key empty": entry("", 0, lp(b"x")),
        "key overlong": entry("", 0, lp(b"x"), raw_key=b"\xc1\xbf"), "key unicode": entry("ключ", 0, lp(b"x")),
        "np ok": entry("a_shape", 5, np_payload()), "np f4": entry("a", 5, np_payload(b"<f4", (3,), np.arange(3, dtype="<f4").tobytes())),
        "np u1": entry("a", 5, np_payload(b"|u1", (3,), b"abc")), "np big endian": entry("a", 5, np_payload(b">i4", (2,))),
        "np native": entry("a", 5, np_payload(b"=i8", (1,), bytes(8))), "np bool": entry("a", 5, np_payload(b"|b1", (2,), b"\x00\x01")),
        "np f2": entry("a", 5, np_payload(b"<f2", (2,), bytes(4))), "np c8": entry("a", 5, np_payload(b"<c8", (1,), bytes(8))),
        "np name float32": entry("a", 5, np_payload(b"float32", (1,), bytes(4))), "np name uint8": entry("a", 5, np_payload(b"uint8", (2,), bytes(2))),
        "np bad dtype": entry("a", 5, np_payload(b"<z9", (2,))), "np empty dtype": entry("a", 5, np_payload(b"", (2,))),
        "np dtype not utf8": entry("a", 5, np_payload(b"\xff4", (2,))), "np f5": entry("a", 5, np_payload(b"<f5", (2,))),
        "np shape mismatch": entry("a", 5, np_payload(b"<i4", (3,), bytes(8))), "np ragged bytes": entry("a", 5, np_payload(b"<i4", (2,), bytes(7))),
        "np 0-d": entry("a", 5, np_payload(b"<i4", (), bytes(4))), "np 0-d wrong": entry("a", 5, np_payload(b"<i4", (), bytes(8))),
        "np empty": entry("a", 5, np_payload(b"<i4", (0,), b"")), "np 0x5": entry("a", 5, np_payload(b"<i4", (0, 5), b"")),
        "np 2x3": entry("a", 5, np_payload(b"<i4", (2, 3))), "np huge shape": entry("a", 5, np_payload(b"<i4", (65536, 65536, 65536), bytes(8))),
        "np data length lies": entry("a", 5, np_payload(b"<i4", (2,), bytes(8), data_len=999)),
        "np data length cuts": entry("a", 5, np_payload(b"<i4", (1,), bytes(8), data_len=4)),
        "np ndim lies": entry("a", 5, np_payload(b"<i4", (2,), ndim=9)), "np empty payload": entry("a", 5, b""),
        "np only dtype length": entry("a", 5, b"\x03"), "np dtype length lies": entry("a", 5, b"\xf0<i4"),
    }
    for name, e in others.items():
        add("entry before: " + name, raw_file([e, E_STREAM, E_META]))
        add("entry after: " + name, raw_file([E_STREAM, E_META, e]))

    # ---- structure: duplicates, counts, lengths that lie, the end of the file ---------------------------------
    bad_meta = entry("clip_meta", 4, lp(b'{"dim": 7}'))
    bad_stream = entry("clip_stream", 0, lp(b"not a frame"))
    add("duplicate meta, last good", raw_file([E_STREAM, bad_meta, E_META]))
    add("duplicate meta, last bad", raw_file([E_STREAM, E_META, bad_meta]))
    add("duplicate stream, last good", raw_file([bad_stream, E_META, E_STREAM]))
    add("duplicate stream, last bad", raw_file([E_STREAM, E_META, bad_stream]))
    add("meta before stream", raw_file([E_META, E_STREAM]))
    add("n_items too small", raw_file([E_STREAM, E_META], n_items=1))
    add("n_items too small, rest is junk", raw_file([E_STREAM, E_META, b"\xff" * 9], n_items=2))
    add("n_items too large", raw_file([E_STREAM, E_META], n_items=3))
    add("n_items huge", raw_file([E_STREAM, E_META], n_items=0xFFFFFFFF))
    add("n_items zero", raw_file([E_STREAM, E_META], n_items=0))
    good = raw_file([E_STREAM, E_META])
    add("trailing junk", good + b"\x00" * 7)
    add("trailing second file", good + good)
    big_tail = raw_file([E_META, E_STREAM, entry("z_bit_stream", 0, lp(bytes(range(200))))])
    for cut in (1, 2, 3, 4, 5, 50, 199, 203, 204, 207, 208, 209, 212):
        add(f"last BYTES entry cut by {cut}", big_tail[:-cut])
    str_tail = raw_file([E_META, E_STREAM, entry("s", 1, lp("é".encode() * 20))])
    for cut in (1, 2, 39, 40, 41, 43, 44, 45):
        add(f"last STR entry cut by {cut}", str_tail[:-cut])
    json_tail = raw_file([E_STREAM, entry("j", 4, lp(b"[1, 2, 3]")), E_META])
    for cut in range(1, len(E_META) + 3, 7):
        add(f"file ending in clip_meta cut by {cut}", json_tail[:-cut])
    for t, p in ((2, bytes(8)), (3, bytes(8)), (7, b"\x01"), (6, b"")):
        tail = raw_file([E_STREAM, E_META, entry("x", t, p)])
        for cut in range(0, len(p) + 2):
            add(f"last fixed-size entry (type {t}) cut by {cut}", tail[:len(tail) - cut])
    np_tail = raw_file([E_STREAM, E_META, entry("a_shape", 5, np_payload(b"<i4", (2, 2)))])
    for cut in range(1, 40, 3):
        add(f"last NP entry cut by {cut}", np_tail[:-cut])
    lying = bytearray(raw_file([E_STREAM, E_META, entry("b", 0, lp(b"abcdef"))]))
    pos = len(lying) - 6 - 4 - 4
    lying[pos:pos + 4] = struct.pack("<I", 1000)        # outer length of the LAST entry beyond the end of the file
    add("last entry's outer length lies", lying)
    lying = bytearray(raw_file([entry("b", 0, lp(b"abcdef")), E_STREAM, E_META]))
    pos = 4 + 2 + 4 + 14 + 4 + 2 + 1 + 1
    lying[pos:pos + 4] = struct.pack("<I", 1000)        # ... of the FIRST entry
    add("first entry's outer length lies", lying)
    for n in (0, 1, 3, 4, 5, 6, 9, 10, 13, 14, 15, 18):
        add(f"file of {n} bytes", good[:n])
    add("bad magic", b"C2DG" + good[4:])
    add("magic lower case", b"c2df" + good[4:])

    # ---- random damage ------------------------------------------------------------------------------------
    def region(b, needle):
        i = b.find(needle)
        return (i, i + len(needle)) if i >= 0 else (0, len(b))

    for s_i, s in enumerate(seeds + [good]):
        hdr_lo, hdr_hi = 10, 10 + struct.unpack_from("<I", s, 6)[0]
        meta_lo, meta_hi = region(s, b'{"model_id"') if s_i == 3 else (s.rfind(b'{"model_id"'), len(s))
        for it in range(260):
            b = bytearray(s)
            kind = it % 5
            if kind == 0:      # flips anywhere
                for _ in range(int(rng.integers(1, 4))):
                    b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:    # flips inside the header JSON
                for _ in range(int(rng.integers(1, 3))):
                    b[int(rng.integers(hdr_lo, max(hdr_hi, hdr_lo + 1)))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 2:    # flips inside clip_meta's JSON
                for _ in range(int(rng.integers(1, 3))):
                    b[int(rng.integers(meta_lo, meta_hi))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 3:    # a byte replaced by a JSON-significant one
                lo, hi = (hdr_lo, hdr_hi) if it % 2 else (meta_lo, meta_hi)
                b[int(rng.integers(lo, max(hi, lo + 1)))] = rng.choice(list(b'"\\{}[],:\x00\x1f\x7f\x80\xc3\xe2\xff 0-9eE.+\t\n'))
            else:              # truncation
                b = b[:int(rng.integers(0, len(b)))]
            add(f"random seed{s_i} kind{kind} #{it}", b)

    # ---- token soup: random sequences of JSON tokens, as clip_meta, as the value of a member, as "dim" ------------
    tokens = [b"{", b"}", b"[", b"]", b",", b":", b'"dim"', b'"a"', b'""', b"64", b"0", b"-1", b"1.5", b"-1.5e3", b"2E+2",
              b"1.", b".5", b"01", b"-", b"+1", b"1e", b"true", b"false", b"null", b"NaN", b"Infinity", b"-Infinity", b"nul",
              b"True", b" ", b"\n", b"\t", b'"\\u0041"', b'"\\q"', b'"x\ty"', b"\xc3\xa9", b'"\xc3\xa9"', b"\x00", b"/", b"'"]
    for it in range(1500):
        soup = b"".join(tokens[int(k)] for k in rng.integers(0, len(tokens), int(rng.integers(1, 10))))
        kind = it % 3
        if kind == 0:
            add(f"soup meta #{it}", with_meta(soup))
        elif kind == 1:
            add(f"soup member #{it}", with_meta(b'{"dim":64,"a":' + soup + b"}"))
        else:
            add(f"soup dim #{it}", with_meta(b'{"a":1,"dim":' + soup + b"}"))
    scalars = [b"64", b"64.0", b"6.4e1", b"640e-1", b"64.5", b"-64", b'"64"', b'" 64"', b'"6_4"', b'"64 "', b"true", b"null",
               b"[64]", b'"\\u00a064"', b'"64\\u2003"', b'"\\u0085 64"', b'"\\uff16\\uff14"', b'"\\u0666\\u0664"', b"1e2", b"0064",
               b'"0_064"', b'"+0064"', b'"-0"', b'"6\\u00a04"', b'"\\u001c64"']
    for a in scalars:
        add("dim " + a.decode("latin-1"), with_meta(b'{"dim":' + a + b"}"))
        add("dim after junk dim " + a.decode("latin-1"), with_meta(b'{"dim":"x","dim":' + a + b"}"))

    # ---- what the reference does ----------------------------------------------------------------------------
    ok, dims, classes, out_codes = [], [], [], []
    for desc, b in cases:
        try:
            z, _ = mg.ref_search.decode_clip_from_c2df(b)
            enc, _ = mg.ref_fm.unpack_c2df(b)
            q = np.frombuffer(mg.ZstdDecompressor().decompress(enc["clip_stream"]), dtype=np.uint8)
            assert q.size == z.shape[0]
            ok.append(True)
            dims.append(q.size)
            classes.append("OK")
            out_codes.append(q)
        except BaseException as e:  # noqa: BLE001  (build.py:87 catches Exception; RecursionError etc. are Exceptions too)
            ok.append(False)
            dims.append(0)
            classes.append(type(e).__name__)
    out = {
        "blob": np.frombuffer(b"".join(b for _, b in cases), dtype=np.uint8),
        "offsets": np.cumsum([0] + [len(b) for _, b in cases]).astype(np.int64),
        "ok": np.array(ok), "dims": np.array(dims, dtype=np.int32), "classes": np.array(classes),
        "desc": np.array([d for d, _ in cases]),
        "codes": np.concatenate(out_codes).astype(np.uint8) if out_codes else np.zeros(0, np.uint8),
    }
    np.savez_compressed(OUT, **out)
    from collections import Counter
    print("wrote", OUT, len(cases), "files,", int(np.sum(ok)), "kept by the reference;", OUT.stat().st_size, "bytes")
    print(Counter(classes).most_common())


if __name__ == "__main__":
    main()
