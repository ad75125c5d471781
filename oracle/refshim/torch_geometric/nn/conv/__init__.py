"""MessagePassing base of torch-geometric 1.6.1, restated for the only call shape the
reference uses: propagate(edge_index=[2,E], x=[N,F], edge_attr=[E,F], size=None) with
flow='source_to_target' (x_j = x[edge_index[0]], aggregate over edge_index[1])."""
import inspect
import torch
from torch_scatter import scatter


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim
        assert flow == "source_to_target"

    def propagate(self, edge_index, size=None, **kwargs):
        x = kwargs["x"]
        n = x.size(self.node_dim)
        msg_args = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_j"):
                msg_args[name] = kwargs[name[:-2]].index_select(self.node_dim, edge_index[0])
            elif name.endswith("_i"):
                msg_args[name] = kwargs[name[:-2]].index_select(self.node_dim, edge_index[1])
            else:
                msg_args[name] = kwargs.get(name)
        out = self.message(**msg_args)
        out = self.aggregate(out, index=edge_index[1], dim_size=n)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, dim_size=None):
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=self.aggr)

    def update(self, inputs):
        return inputs
