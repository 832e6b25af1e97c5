from .hierarchical_graph import HierarchicalGraph  # noqa: F401
