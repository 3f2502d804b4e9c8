def copy_u(u, out):
    """dgl.function.copy_u: message[out] = source node feature u."""
    def f(edges):
        return {out: edges.src[u]}
    return f


copy_src = copy_u
