"""World-size-2 gloo test (CPU) of the N > 1 host logic: LPT sharding, the single all-reduce, BED gather.

The per-chromosome compute here is the ORACLE (no GPU in this test); what is checked is that two ranks
working on their shards reproduce exactly what one rank does on the whole set."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from oracle import oracle as orc
        from rocco_b200 import distributed as rd
        names, sizes, seeds = _toy_genome()
        mine = rd.shard_chromosomes(sizes)
        recs, sel, bins = [], 0, 0
        for i in mine:
            s = _toy_scores(sizes[i], seeds[i])
            sol, _ = orc.solve_chrom_exact(s, budget=0.05, gamma=1.0)
            recs += orc.solution_to_records(names[i], np.arange(0, 50 * sizes[i], 50), sol)
            sel += int(sol.sum())
            bins += sizes[i]
        tot_sel, tot_bins = rd.allreduce_selected(sel, bins)
        merged = rd.gather_bed_records(recs)
        # the tensor-collective variant: runs as (global chromosome index in record order, first bin, last bin + 1)
        lex = sorted(names)
        runs = rd.gather_runs([lex.index(c) for c, _, _ in recs], [a // 50 for _, a, _ in recs], [b // 50 for _, _, b in recs])
        if rank == 0:
            with open(os.path.join(out_dir, "merged_runs.bed"), "w") as fh:
                fh.write("".join(f"{lex[c]}\t{50 * a}\t{50 * b}\n" for c, a, b in zip(*[r.tolist() for r in runs])))
        else:
            assert runs is None
        # the positioned-write variant: every rank writes its chromosomes straight into ONE file (a stale, longer file is there)
        shared = os.path.join(out_dir, "genome_positioned.bed")
        if rank == 0:
            with open(shared, "wb") as fh:
                fh.write(b"#" * 200000)
        dist.barrier()
        my_names = [names[i] for i in mine]
        local = {c: k for k, c in enumerate(my_names)}
        my_runs = (np.array([local[c] for c, _, _ in recs], dtype=np.int32), np.array([a // 50 for _, a, _ in recs], dtype=np.int64),
                   np.array([b // 50 for _, _, b in recs], dtype=np.int64))
        tot = rd.write_genome_bed(shared, lex, my_names, my_runs, 50, extras=(sel, bins))
        assert tot == [tot_sel, tot_bins]
        dist.barrier()
        np.save(os.path.join(out_dir, f"tot_{rank}.npy"), np.array([tot_sel, tot_bins]))
        if rank == 0:
            with open(os.path.join(out_dir, "merged.bed"), "w") as fh:
                fh.write("".join(f"{c}\t{a}\t{b}\n" for c, a, b in merged))
    finally:
        dist.destroy_process_group()


def _toy_genome():
    names = ["chr1", "chr2", "chr10", "chrX", "chr21"]
    sizes = [5000, 4200, 2600, 3100, 900]
    return names, sizes, [11, 12, 13, 14, 15]


def _toy_scores(n, seed):
    rng = np.random.default_rng(seed)
    return rng.normal(size=n) - 0.8 + 4.0 * (rng.random(n) < 0.03) * rng.random(n)


def test_two_ranks_reproduce_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    from rocco_b200.pipeline import lpt_partition

    names, sizes, seeds = _toy_genome()
    parts = lpt_partition(sizes, 2)
    assert sorted(parts[0] + parts[1]) == list(range(5)) and parts[0] and parts[1]
    assert abs(sum(sizes[i] for i in parts[0]) - sum(sizes[i] for i in parts[1])) <= max(sizes)

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)

    recs, sel = [], 0
    for n, sd, c in zip(sizes, seeds, names):
        sol, _ = orc.solve_chrom_exact(_toy_scores(n, sd), budget=0.05, gamma=1.0)
        recs += orc.solution_to_records(c, np.arange(0, 50 * n, 50), sol)
        sel += int(sol.sum())
    want = "".join(f"{c}\t{a}\t{b}\n" for c, a, b in orc.merge_bed_records(recs))
    assert open(tmp_path / "merged.bed").read() == want
    assert open(tmp_path / "merged_runs.bed").read() == want
    assert open(tmp_path / "genome_positioned.bed").read() == want
    for r in range(2):
        assert np.load(tmp_path / f"tot_{r}.npy").tolist() == [sel, sum(sizes)]


def test_lpt_balance_hg38():
    """balance 0.999 / 0.995 / 0.965 at 2 / 4 / 8 ranks"""
    from rocco_b200.pipeline import lpt_partition
    from rocco_b200.synth import HG38_SIZES, chrom_bins
    bins = [chrom_bins(c) for c in HG38_SIZES]
    for parts, floor in ((2, 0.99), (4, 0.99), (8, 0.96)):
        loads = [sum(bins[i] for i in p) for p in lpt_partition(bins, parts)]
        assert sum(loads) == sum(bins) == 61765409
        assert (sum(loads) / parts) / max(loads) >= floor
