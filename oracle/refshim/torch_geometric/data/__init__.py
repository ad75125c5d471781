import torch


class Data(object):
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class Batch(Data):
    pass


class DataLoader(torch.utils.data.DataLoader):
    pass
