def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
