"""The single-process multi-GPU index (``faiss.IndexFlatIP(d, devices=[...])`` → ``sgic_index_create_sharded``,
SURVEY.md §8e): one handle, one process, rows sharded over several GPUs, answers identical to one GPU.

``devices`` may name a GPU several times, so the whole sharded path — placement of appends, row-number maps,
peer stores into the home gather buffer, the K5 merge, shard files — is exercised on a ONE-GPU box too; with more
GPUs visible the same tests spread the shards over real devices.  The reference's caller is one process
(src/search.py:149-162, webapp.py:246-248); its search call is src/search.py:115.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c2df_ref
from oracle.flat_ip import check_topk


def unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def f16(x):
    return x.astype(np.float16).astype(np.float32)


def device_list(n_shards):
    import torch
    g = torch.cuda.device_count()
    return [i % g for i in range(n_shards)]


@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_sharded_answers_equal_single_gpu_answers(n_shards):
    from sgic_b200 import faiss_compat as faiss
    rng = np.random.default_rng(100 + n_shards)
    n, d = 90_000, 512
    xb = unit(rng, n, d)
    xb[5000:5040] = xb[100:140]              # exact duplicates across shard boundaries: ties resolve by global row
    xb[80_000:80_040] = xb[100:140]
    xq = np.concatenate([xb[100:104] + 0.0, unit(rng, 300, d)])
    one = faiss.IndexFlatIP(d, device=0)
    one.add(xb)
    many = faiss.IndexFlatIP(d, devices=device_list(n_shards))
    assert many.n_shards == n_shards and many.d == d and many.ntotal == 0
    many.add(xb)                              # one large block: n_shards contiguous slices
    assert many.ntotal == n and many.stat("n_segments") == n_shards
    for nq, k in ((1, 10), (2, 10), (7, 5), (64, 10), (304, 10), (33, 100), (1, 1500)):
        D1, I1 = one.search(xq[:nq], k)
        Dm, Im = many.search(xq[:nq], k)
        assert np.array_equal(I1, Im), (nq, k)
        assert np.array_equal(D1, Dm), (nq, k)
    D, I = many.search(xq[:4], 3)
    assert [list(r) for r in I] == [[100 + j, 5000 + j, 80_000 + j] for j in range(4)]
    check_topk(*many.search(xq[:16], 10), f16(xb), f16(xq[:16]), 10, score_tol=3e-5, tie_tol=1e-6)
    # k beyond the rows of a shard / of the index: padding comes out once, at the end
    small = faiss.IndexFlatIP(d, devices=device_list(n_shards))
    small.add(xb[:5])
    D, I = small.search(xq[:2], 8)
    assert np.all(I[:, 5:] == -1) and np.all(I[:, :5] >= 0) and np.all(D[:, 5:] < -3e38)
    for idx in (one, many, small):
        idx.close()


def test_launch_threads_and_single_thread_give_the_same_answer():
    from sgic_b200 import faiss_compat as faiss
    rng = np.random.default_rng(7)
    xb, xq = unit(rng, 40_000, 256), unit(rng, 20, 256)
    idx = faiss.IndexFlatIP(256, devices=device_list(4))
    idx.add(xb)
    ref = idx.search(xq, 10)
    idx.set_option("workers", 0)
    got = idx.search(xq, 10)
    assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])
    idx.set_option("workers", 1)
    for _ in range(50):                       # back-to-back searches reuse the gather buffer and the events
        got = idx.search(xq[:1], 10)
        assert np.array_equal(got[1], ref[1][:1])
    idx.close()


def test_small_appends_balance_and_row_numbers_stay_global():
    """compress.py:300-305 adds vectors one at a time; blocks of every size may follow each other.  Rows keep the
    order they were added in (ids.txt line i <-> row i), whatever GPU they landed on."""
    from sgic_b200 import faiss_compat as faiss
    rng = np.random.default_rng(11)
    d = 128
    xb = unit(rng, 200_000, d)
    many = faiss.IndexFlatIP(d, devices=device_list(3))
    pos = 0
    for blk in (1, 1, 1, 70_000, 3, 5000, 1, 66_000, 2, 40_000, 1000, 17_990, 1):
        many.add(xb[pos:pos + blk])
        pos += blk
    assert many.ntotal == pos == 200_000
    assert many.stat("n_segments") > 3        # several runs per GPU: the remap kernel is on the path
    sizes = [many.shard(g).ntotal for g in range(3)]
    assert min(sizes) > 20_000, sizes         # nobody starved
    one = faiss.IndexFlatIP(d, device=0)
    one.add(xb)
    xq = unit(rng, 40, d)
    for nq, k in ((1, 10), (40, 10), (9, 64)):
        D1, I1 = one.search(xq[:nq], k)
        Dm, Im = many.search(xq[:nq], k)
        assert np.array_equal(I1, Im) and np.array_equal(D1, Dm)
    # rows come back by GLOBAL number
    for i0, cnt in ((0, 5), (69_999, 10), (141_000, 3000), (199_990, 10)):
        assert np.array_equal(many.reconstruct_n(i0, cnt), one.reconstruct_n(i0, cnt))
    one.close()
    many.close()


def test_write_index_and_codes_through_the_front(tmp_path):
    from sgic_b200 import c2df as c2, faiss_compat as faiss, zstd
    from sgic_b200.index_build import quantize_u8_and_compress
    rng = np.random.default_rng(13)
    d, n = 64, 30_000
    xb = unit(rng, n, d)
    # fp32 rows retained on every shard -> the IxFI file equals the single-GPU one (and what faiss would write)
    one = faiss.IndexFlatIP(d, device=0)
    one.add(xb)
    many = faiss.IndexFlatIP(d, devices=device_list(3))
    many.add(xb[:20_000])
    many.add(xb[20_000:20_007])
    many.add(xb[20_007:])
    faiss.write_index(one, str(tmp_path / "one.index"))
    faiss.write_index(many, str(tmp_path / "many.index"))
    assert (tmp_path / "one.index").read_bytes() == (tmp_path / "many.index").read_bytes()
    back = faiss.read_index(str(tmp_path / "many.index"), device=0)
    assert np.array_equal(back.reconstruct_n(0, n), one.reconstruct_n(0, n))
    # .c2df ingest over the shards: file order = row order, failed files skipped, codes retained -> byte-identical IxFI
    blobs, codes = [], []
    for i, v in enumerate(xb[:9000]):
        stream, meta = quantize_u8_and_compress(v)
        if i in (17, 3000, 3001, 8999):
            blobs.append(b"C2DFbroken")
            continue
        codes.append(np.frombuffer(zstd.decompress(stream), dtype=np.uint8))
        blobs.append(c2.pack_c2df({"clip_stream": stream, "clip_meta": meta}, {"version": 2}))
    offs = np.zeros(len(blobs) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in blobs], out=offs[1:])
    ing = faiss.IndexFlatIP(d, devices=device_list(3), retain_fp32=False, retain_codes=True)
    added, status = ing.add_c2df(np.frombuffer(b"".join(blobs), dtype=np.uint8), offs)
    assert added == 8996 and [i for i, s in enumerate(status) if s] == [17, 3000, 3001, 8999]
    codes = np.stack(codes)
    assert np.array_equal(ing.codes(), codes)
    faiss.write_index(ing, str(tmp_path / "ing.index"))
    c2df_ref.write_ixfi(tmp_path / "ref.index", np.stack([c2df_ref.dequantize_clip_u8(q) for q in codes]))
    assert (tmp_path / "ing.index").read_bytes() == (tmp_path / "ref.index").read_bytes()
    for idx in (one, many, back, ing):
        idx.close()


def test_shard_files_round_trip_and_direct_shard_fill(tmp_path):
    import torch
    from sgic_b200 import faiss_compat as faiss
    from sgic_b200.synth import fill_index_random, random_unit_queries
    d, per = 512, 120_000
    devs = device_list(3)
    many = faiss.IndexFlatIP(d, devices=devs, retain_fp32=False)
    for g in range(3):                        # bulk load: every shard filled on its own GPU, then adopted
        fill_index_random(many.shard(g), per, row0=g * per, chunk_rows=per)
    many.adopt_shards()
    assert many.ntotal == 3 * per
    one = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
    for g in range(3):
        fill_index_random(one, per, row0=g * per, chunk_rows=per)
    xq = random_unit_queries(70, d)
    D1, I1 = one.search(xq, 10)
    Dm, Im = many.search(xq, 10)
    assert np.array_equal(I1, Im) and np.array_equal(D1, Dm)
    many.save_shards(tmp_path / "shards")
    assert sorted(p.name for p in (tmp_path / "shards").iterdir()) == [f"shard-{g:05d}-of-00003.sgi2" for g in range(3)]
    back = faiss.read_index_shards(tmp_path / "shards", devs)
    assert back.ntotal == 3 * per and back.n_shards == 3
    Db, Ib = back.search(xq, 10)
    assert np.array_equal(Ib, I1) and np.array_equal(Db, D1)
    # device-resident call: queries and answers on the home GPU
    home = torch.device("cuda", devs[0])
    qd = torch.from_numpy(xq).to(home)
    Dd, Id = back.search_torch(qd, 10)
    torch.cuda.synchronize(home)
    assert np.array_equal(Id.cpu().numpy(), I1) and np.array_equal(Dd.cpu().numpy(), D1)
    # appends from a tensor on the home GPU are cut over the shards as well
    extra = torch.from_numpy(f16(unit(np.random.default_rng(3), 50_000, d))).to(home)
    back.add_torch(extra)
    one.add_torch(extra.to(torch.device("cuda", 0)))
    Db, Ib = back.search(xq[:5], 10)
    D1, I1 = one.search(xq[:5], 10)
    assert back.ntotal == one.ntotal and np.array_equal(Ib, I1) and np.array_equal(Db, D1)
    with pytest.raises(RuntimeError, match="sgic_index_save_shards"):
        faiss.write_shard(back, str(tmp_path / "x.sgi2"))
    for idx in (one, many, back):
        idx.close()


def test_service_runs_on_a_multi_gpu_index(tmp_path):
    """N2 on all GPUs: the resident service over a sharded directory gives the single-GPU answers."""
    import json
    from sgic_b200 import faiss_compat as faiss, retrieval
    from sgic_b200.service import SearchService
    rng = np.random.default_rng(21)
    d, n = 512, 60_000
    xb = unit(rng, n, d)
    paths = [f"../IO/bitstreams/img{i:06d}.c2df" for i in range(n)]
    many = faiss.IndexFlatIP(d, devices=device_list(4), retain_fp32=False)
    many.add(xb)
    many.save_shards(tmp_path)
    (tmp_path / "paths.json").write_text(json.dumps(paths), encoding="utf-8")
    (tmp_path / "meta.json").write_text(json.dumps({"dim": d, "model_id": None}), encoding="utf-8")
    index, got_paths, meta = retrieval.load_index(tmp_path, devices=device_list(4))
    assert index.n_shards == 4 and index.ntotal == n and got_paths == paths and meta["dim"] == d
    svc = SearchService(index=index, paths=got_paths, meta=meta)
    res = svc.search_vec(xb[4242], topk=5)
    assert res[0][0] == paths[4242] and abs(res[0][1] - 1.0) < 2e-3 and len(res) == 5
    one = faiss.IndexFlatIP(d, device=0)
    one.add(xb)
    want = retrieval.do_search(xb[4242:4243], one, paths, topk=5)
    assert [p for p, _ in res] == [p for p, _ in want]
    svc.close()
    for idx in (one, many, index):
        idx.close()


def test_failed_append_leaves_no_rows_behind(tmp_path):
    """An append that fails on ONE shard (out of memory, here injected) is undone on the others: the index keeps its
    row count, its row numbers and its answers, and the same append succeeds afterwards."""
    import torch
    from sgic_b200 import c2df as c2, faiss_compat as faiss
    from sgic_b200.index_build import quantize_u8_and_compress
    rng = np.random.default_rng(31)
    d = 64
    xb = unit(rng, 60_000, d)
    many = faiss.IndexFlatIP(d, devices=device_list(3), retain_codes=True)
    one = faiss.IndexFlatIP(d, device=0)
    many.add(xb[:20_000])
    one.add(xb[:20_000])
    xq = unit(rng, 9, d)
    want = one.search(xq, 10)
    # host rows, cut over the three shards: the middle one fails
    many.shard(1).set_option("fail_appends", 1)
    with pytest.raises(RuntimeError, match="shard 1"):
        many.add(xb[20_000:50_000])
    assert many.ntotal == 20_000 and [many.shard(g).ntotal for g in range(3)] == [one.ntotal // 3 + (g < one.ntotal % 3) for g in range(3)]
    got = many.search(xq, 10)
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])
    # device rows on the home GPU: the last shard fails after the first two have taken their slices
    home = torch.device("cuda", many.shard(0).device)
    many.shard(2).set_option("fail_appends", 1)
    with pytest.raises(RuntimeError):
        many.add_torch(torch.from_numpy(xb[20_000:50_000]).to(home))
    assert many.ntotal == 20_000
    # .c2df files: shard 0 fails
    blobs = []
    for v in xb[50_000:56_000]:
        stream, meta = quantize_u8_and_compress(v)
        blobs.append(c2.pack_c2df({"clip_stream": stream, "clip_meta": meta}, {"version": 2}))
    offs = np.zeros(len(blobs) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in blobs], out=offs[1:])
    blob = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    many.shard(0).set_option("fail_appends", 1)
    with pytest.raises(RuntimeError):
        many.add_c2df(blob, offs)
    assert many.ntotal == 20_000
    got = many.search(xq, 10)
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])
    # the same appends go through afterwards, and the row numbers are the single-GPU ones
    many.add(xb[20_000:50_000])
    one.add(xb[20_000:50_000])
    added, status = many.add_c2df(blob, offs)
    assert added == 6000 and not status.any() and many.ntotal == 56_000
    one.add(many.reconstruct_n(50_000, 6000))
    for nq, k in ((1, 10), (9, 10)):
        D1, I1 = one.search(xq[:nq], k)
        Dm, Im = many.search(xq[:nq], k)
        assert np.array_equal(I1, Im) and np.array_equal(D1, Dm)
    one.close()
    many.close()
