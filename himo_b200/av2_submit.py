"""AV2 scene-flow leaderboard submission files (the wire format after the path in `data_mode=test`):
OSF/src/utils/av2_eval.py:758-801 (`write_output_file`) and OSF/src/utils/mics.py:312-344 (`zip_res`).

  version 1 (eval.ai 2010): per sweep `<log>/<timestamp>.feather` with float16 `flow_t{x,y,z}_m` of the evaluated
                            points (total flow) and a bool `is_dynamic` (|flow - rigid flow| >= 0.05 m, trainer.py:284)
  version 2 (eval.ai 2210): float16 flow RELATIVE to the ego motion for ALL points plus `is_valid` (all true); inside
                            the zip the sweeps of a log are renamed 0000000000, 0000000005, ... in timestamp order and a
                            `metadata.json` states whether labels were used
"""
from __future__ import annotations

import json
import os
import shutil
from pathlib import Path
from typing import Tuple
from zipfile import ZipFile

import numpy as np

DYNAMIC_THRESHOLD_M = 0.05


def leaderboard_arrays(final_flow: np.ndarray, pose_flow: np.ndarray, eval_mask: np.ndarray, leaderboard_version: int):
    """trainer.py:281-288: what is written for one sweep -> (flow [K,3] float32, is_dynamic [M] bool)."""
    pred = final_flow[eval_mask, :3]
    rigid = pose_flow[eval_mask, :3]
    is_dynamic = np.linalg.norm(pred - rigid, axis=1, ord=2) >= DYNAMIC_THRESHOLD_M
    if leaderboard_version == 2:
        pred = final_flow - pose_flow            # every point, ego motion removed
    return pred, is_dynamic


def write_output_file(flow: np.ndarray, is_dynamic: np.ndarray, sweep_uuid: Tuple[str, int], output_dir,
                      leaderboard_version: int = 1) -> None:
    import pandas as pd
    if leaderboard_version not in (1, 2):
        raise ValueError(f"Leaderboard version {leaderboard_version} is not valid. Please set it to 1 or 2.")
    log_dir = Path(output_dir) / str(sweep_uuid[0])
    log_dir.mkdir(exist_ok=True, parents=True)
    cols = {f"flow_t{a}_m": np.asarray(flow)[:, i].astype(np.float16) for i, a in enumerate("xyz")}
    if leaderboard_version == 1:
        cols["is_dynamic"] = np.asarray(is_dynamic).astype(bool)
    else:
        cols = {"is_valid": np.ones(np.asarray(flow).shape[0], dtype=bool), **cols}
    pd.DataFrame(cols).to_feather(log_dir / f"{sweep_uuid[1]}.feather")


def zip_res(res_folder, output_file: str = "av2_submit.zip", leaderboard_version: int = 2, is_supervised: bool = False,
            remove_folder: bool = False) -> str:
    res_folder = str(res_folder)
    logs = sorted(d for d in os.listdir(res_folder) if os.path.isdir(os.path.join(res_folder, d)))
    if leaderboard_version != 1 and output_file == "av2_submit.zip":
        output_file = output_file.replace(".zip", f"_v{leaderboard_version}.zip")
    with ZipFile(output_file, "w") as z:
        if leaderboard_version != 1:
            z.writestr("metadata.json", json.dumps({"Is Supervised?": is_supervised}, indent=4))
        for log in logs:
            sweeps = sorted(f for f in os.listdir(os.path.join(res_folder, log)) if f.endswith(".feather"))
            for k, name in enumerate(sweeps):
                arc = name if leaderboard_version == 1 else f"{5 * k:010d}.feather"
                z.write(os.path.join(res_folder, log, name), arcname=os.path.join(log, arc))
    if remove_folder:
        shutil.rmtree(res_folder)
    return output_file
