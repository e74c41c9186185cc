"""Overlay mechanics: this `src` package shadows the reference's `src` for the modules on the scoring path and
falls through to the reference tree (env VITAD_REFERENCE_ROOT, default: not set) for everything else
(data loaders, learners, plotting), so the reference's scripts keep importing `src.…` unchanged."""
import os


def extend(path_list, *sub):
    root = os.environ.get("VITAD_REFERENCE_ROOT")
    if root:
        cand = os.path.join(root, "src", *sub)
        if os.path.isdir(cand) and cand not in path_list:
            path_list.append(cand)
    return path_list
