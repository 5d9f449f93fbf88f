from xmap_b200.core.recommender import RecommenderSim  # noqa: F401
