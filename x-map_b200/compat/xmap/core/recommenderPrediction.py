from xmap_b200.core.recommender import RecommenderPrediction  # noqa: F401
