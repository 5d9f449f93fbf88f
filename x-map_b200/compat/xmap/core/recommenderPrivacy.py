from xmap_b200.core.recommender import RecommenderPrivacy  # noqa: F401
