"""Host-side multi-rank logic on CPU (gloo, world_size 2): sharding, record packing, the gather that replaces
collect_results_cpu, and evaluate() giving the same numbers as a single process."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import tiseg_b200  # noqa: E402,F401
from tiseg_b200 import datasets, parallel, synth  # noqa: E402
from refpipe import oracle_pre_eval, same_result  # noqa: E402


def _make(n, C, multi):
    tiles = [synth.gt_and_pred(4000 + i, 64, 72, num_classes=C) for i in range(n)]
    res = [oracle_pre_eval(dict(sem_pred=t['pred_sem'], inst_pred=t['pred_inst']), t['gt_sem'], t['gt_inst'], C,
                           multi=multi, name=None if multi else "img%d" % i) for i, t in enumerate(tiles)]
    return res


def test_pack_unpack_roundtrip():
    for C, multi in ((2, False), (7, True)):
        res = _make(3, C, multi)
        back = parallel.unpack_results(parallel.pack_results(res, C), C, None if multi else [r['name'] for r in res])
        for a, b in zip(res, back):
            same_result(a, b)


def test_shard_indices_interleave():
    assert parallel.shard_indices(7, 0, 2) == [0, 2, 4, 6] and parallel.shard_indices(7, 1, 2) == [1, 3, 5]
    assert sorted(sum((parallel.shard_indices(4981, r, 8) for r in range(8)), [])) == list(range(4981))


def _worker(rank, world, port, C, multi, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = _make(n, C, multi)
        mine = parallel.shard_indices(n, rank, world)
        names = None if multi else [r['name'] for r in res]
        full = parallel.gather_results([res[i] for i in mine], mine, n, C, names)
        ints = parallel.all_reduce_sums([sum(r['bin_pq_pre_eval_res'][0] for r in (res[i] for i in mine)), len(mine)])
        if rank == 0:
            for a, b in zip(res, full):
                same_result(a, b)
            ds = (datasets.CoNICDataset if multi else datasets.CustomDataset)(inst_gts=[None] * n, sem_gts=[None] * n)
            want, _ = ds.evaluate(res, logger="silent")
            got, _ = ds.evaluate(full, logger="silent")
            assert want == got and len(want) > 8
            assert int(ints[0]) == sum(r['bin_pq_pre_eval_res'][0] for r in res) and int(ints[1]) == n
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("C,multi", [(2, False), (7, True)])
def test_gather_results_gloo_world2(C, multi):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (1 if multi else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, C, multi, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(msg == "ok" for _, msg in out), out


def test_gt_prefetch_reads_the_reference_file_layout(tmp_path):
    """custom.py:252-259: *_sem.png through pillow, *_inst.npy through np.load; the prefetcher must hand pre_eval
    exactly what a synchronous read gives, in any order of use."""
    from PIL import Image
    from tiseg_b200 import datasets
    rng = np.random.default_rng(9)
    names = ["img%02d" % i for i in range(7)]
    want = {}
    for nm in names:
        sem = rng.integers(0, 2, (20, 31)).astype(np.uint8)
        inst = rng.integers(0, 50, (20, 31)).astype(np.int32)
        Image.fromarray(sem).save(str(tmp_path / (nm + "_sem.png")))
        np.save(str(tmp_path / (nm + "_inst.npy")), inst)
        want[nm] = (sem, inst)
    ds = datasets.CustomDataset(img_dir=str(tmp_path), ann_dir=str(tmp_path))
    assert len(ds) == 7
    ds.prefetch([5, 1, 3], workers=2)
    for i in (3, 0, 5, 1, 6):                      # prefetched and not, out of order
        sem, inst = ds._load_gt(i)
        nm = os.path.basename(ds.data_infos[i]['sem_file_name'])[:-len('_sem.png')]
        assert np.array_equal(sem, want[nm][0]) and np.array_equal(inst, want[nm][1])
    assert not ds._pending


def _records_worker(rank, world, port, n, width, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch
        full = np.random.default_rng(3).integers(0, 1 << 40, size=(n, width)).astype(np.float64) + 0.25
        mine = parallel.shard_indices(n, rank, world)
        got = parallel.gather_records(torch.from_numpy(full[mine]), mine, n)
        assert got.dtype == np.float64 and np.array_equal(got, full)
        if rank == 1:                       # a rank that lost a tile is an error on every rank, not a silent zero row
            mine = mine[:-1]
        try:
            parallel.gather_records(torch.from_numpy(full[mine]), mine, n)
            raise AssertionError("missing record not detected")
        except RuntimeError as e:
            assert "no record" in str(e)
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_gather_records_gloo_world2():
    """records kept as one [n, R] tensor per rank (Dataset.pre_eval_records): one all-gather, rows at dataset index"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 7
    procs = [ctx.Process(target=_records_worker, args=(r, 2, port, 7, 78, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for _, msg in got), got
