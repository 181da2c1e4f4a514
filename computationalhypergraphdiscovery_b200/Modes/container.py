"""Cluster partition of the variables -> 0/1 index matrix (C x N).
Mirror of the reference's Modes/container.py (CONT:21-42 ctor, :44-66 assign_clusters,
:68-96 make_index_matrix, :98-107 repr), on NumPy."""
import numpy as np


class ModeContainer:
    def __init__(self, variable_names, clusters=None) -> None:
        assert len(list(variable_names)) == len(set(list(variable_names)))
        self.names = variable_names
        self.name_to_index = {name: i for i, name in enumerate(self.names)}
        self.assign_clusters(clusters)
        self.index_matrix = ModeContainer.make_index_matrix(self.names, self.clusters, self.name_to_index)

    def assign_clusters(self, clusters):
        if clusters is None:
            self.clusters = [[name] for name in self.names]
        else:
            clusters_list = [name for cluster in clusters for name in cluster]
            assert len(clusters_list) == len(set(clusters_list))
            assert len(clusters_list) == len(self.names)
            assert set(clusters_list).issubset(set(list(self.names)))
            self.clusters = clusters

    def make_index_matrix(names, clusters, name_to_index):
        res = np.zeros((len(clusters), len(names)))
        for i, cluster in enumerate(clusters):
            for name in cluster:
                if name in name_to_index:
                    res[i, name_to_index[name]] = 1.0
        return res

    @property
    def cluster_of(self):
        """cluster index of every variable (int32, N)."""
        return np.argmax(self.index_matrix, axis=0).astype(np.int32)

    def __repr__(self) -> str:
        res = ["/".join(cluster) for cluster in self.clusters]
        return res.__repr__()
