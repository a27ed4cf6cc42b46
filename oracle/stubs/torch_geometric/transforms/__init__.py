from oracle.pyg import LaplacianLambdaMax  # noqa: F401
