def apply_freeu(resolution_idx, hidden_states, res_hidden_states, **kw):
    raise NotImplementedError("FreeU is never enabled on the Uni-Renderer hot path")
