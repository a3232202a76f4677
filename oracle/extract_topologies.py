"""Extract the real road-graph topologies shipped with the reference
(datasets/G_nwb/4.GEPHI_to_SIM/*.bin) into a small npz the GPU box can read.

Follows the reference's own construction: W = nx.to_numpy_matrix(G) in G.nodes()
order (modules/sim/simdata_utils.py:513; labels == matrix order for these files)
and EI = dense_to_sparse(W)[0] = row-major non-zeros (simdata_utils.py:519).
Run in the build container only:  python -m oracle.extract_topologies
"""
import os
import pickle

import networkx as nx
import numpy as np

from oracle.ref_import import REFERENCE_ROOT

FILES = {
    "NWB_AMS": "G_test_DAM_1km_edited_V=975.bin",
    "NWB_UTR": "G_test_UTR_1km_edited_V=1182.bin",
    "NWB_ROT": "G_test_ROT_2km_edited_V=2602.bin",
}


def main():
    out = {}
    for name, fname in FILES.items():
        with open(os.path.join(REFERENCE_ROOT, "datasets/G_nwb/4.GEPHI_to_SIM", fname), "rb") as f:
            d = pickle.load(f)
        G = d["G"]
        nodes = list(G.nodes())
        assert [d["labels"][n] for n in nodes] == list(range(len(nodes)))
        W = nx.to_numpy_array(G)
        assert set(np.unique(W)) <= {0.0, 1.0}
        ei = np.stack(np.nonzero(W)).astype(np.int32)          # row-major == dense_to_sparse
        out[name + "/edge_index"] = ei
        out[name + "/num_nodes"] = np.int32(len(nodes))
        out[name + "/target_nodes"] = np.array(sorted(d["target_nodes"]), dtype=np.int32)
        print(name, len(nodes), ei.shape[1], len(d["target_nodes"]))
    dst = os.path.join(os.path.dirname(__file__), "..", "simmodel_b200", "data", "nwb_topologies.npz")
    np.savez_compressed(dst, **out)
    print("wrote", os.path.abspath(dst), os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
