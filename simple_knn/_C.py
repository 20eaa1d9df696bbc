from r2_gaussian_b200.simple_knn import distCUDA2  # noqa: F401

__all__ = ["distCUDA2"]
