"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), origin-sharded run of the real engine.

Every rank runs ``ShardedDynamics`` around its own ``DynamicsEngine``; rank 0 also runs the same frames on
a single-GPU engine and through the CPU oracle and writes the comparison to the JSON file given as argv[1].
"""

import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
for p in (REPO, REPO / "statdyn-analysis_b200", REPO / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    out_path = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4099
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)

    from _helpers import WAVE_NUMBER, compare_rows, oracle_process
    from sdanalysis_b200.distributed import ShardedDynamics, partition
    from sdanalysis_b200.engine import DynamicsEngine
    from sdanalysis_b200.molecules import Trimer
    from sdanalysis_b200.synthetic import box_for, linear_schedule, synthetic_frames

    schedule = list(linear_schedule(40, 3, 13))  # 14 origins, created every 3 frames
    frames = list(synthetic_frames(schedule, n, seed=17, sigma_r=0.06, sigma_theta=0.15))  # same on every rank
    report = {"n": n, "world": world}

    def sharded_run(mode):
        engine = DynamicsEngine(n, box_for(n), WAVE_NUMBER, mol_sites=Trimer().positions, device=device)

        def step_fn(ts, pos, quat, keys, rows_out=0):
            return engine.step(ts, pos, quat, keys, to_host=False, rows_out=rows_out)

        sharded = ShardedDynamics(n, step_fn, max_local=16, device=device, batch=16, direct_rows=True)
        rows = []
        for indexes, frame in frames:
            if mode == "sliced":  # every rank holds the frame in host memory and uploads its slice
                handle = sharded.post_frame_sliced(frame.position, frame.orientation)
            else:  # rank 0 is the reader
                handle = sharded.post_frame(*((frame.position, frame.orientation) if rank == 0 else (None, None)))
            got = sharded.step(frame.timestep, list(indexes), handle, gather="batch")
            if got:
                rows.extend(got)
        rows.extend(sharded.flush())
        torch.cuda.synchronize()
        mine = partition(list(range(14)), rank, world)
        state = {k: (engine.delta_arrays(k), engine.summary(k)) for k in mine}
        return sharded, engine, rows, state

    results = {}
    for mode in ("reader", "sliced"):
        sharded, engine, rows, state = sharded_run(mode)
        gathered = [None] * world
        dist.gather_object(state, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            merged = {}
            for part in gathered:
                merged.update(part)
            table = []
            for keys, timediffs, sums in rows:
                for row, key in zip(engine.finalize_rows(sums, timediffs), keys):
                    row["keyframe"] = key
                    table.append(row)
            results[mode] = (table, merged, sharded.exchange, sharded.exchange_note)
        del sharded, engine
        dist.barrier()

    if rank == 0:
        # single GPU, same frames
        from sdanalysis_b200.read import process_frames

        single_rows = []
        single = process_frames(iter(frames), WAVE_NUMBER, sink=single_rows.append, device=device)
        ref_rows, summaries, near, keyframes = oracle_process(frames)
        order = lambda rows: {(r["keyframe"], r["time"]): r for r in rows}  # noqa: E731
        s_rows = order(single_rows)
        for mode, (table, merged, exchange, note) in results.items():
            m_rows = order(table)
            assert set(m_rows) == set(s_rows), "row sets differ"
            identical = all(
                all((m_rows[k][q] == s_rows[k][q]) or (np.isnan(m_rows[k][q]) and np.isnan(s_rows[k][q])) for q in m_rows[k])
                for k in s_rows
            )
            compare_rows([m_rows[(r["keyframe"], r["time"])] for r in ref_rows], ref_rows, n)
            state_identical, passage_mismatch = True, 0
            for key, (dyn, _) in keyframes.items():
                (trans, rot), summary = merged[key]
                s_trans, s_rot = single.delta_arrays(key)
                state_identical &= bool(np.array_equal(trans, s_trans) and np.array_equal(rot, s_rot))
                assert np.array_equal(trans, dyn.delta_translation), f"delta_translation of origin {key}"
                for name, want in summaries[key].items():
                    mask = ~near[key][name]
                    passage_mismatch += int((summary[name][mask] != want[mask]).sum())
                    assert np.array_equal(summary[name], single.summary(key)[name]), f"{name} of origin {key} vs single GPU"
            report[mode] = {
                "rows": len(table), "rows_identical_to_single_gpu": bool(identical), "state_identical_to_single_gpu": state_identical,
                "rows_match_oracle": True, "passage_mismatch_vs_oracle": passage_mismatch, "exchange": exchange, "note": note,
            }
        Path(out_path).write_text(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
