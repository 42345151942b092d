"""Host mirror of the reference's teacher-student variant (code/teacherstudent.py) on the walk kernels.

    CRWBase            :11-268   the CRW module with a softmax transition matrix (:66-80) and a biased Linear head (:43-53)
    SoftCrossEntropyLoss :270-292  cross-entropy between two probability tensors
    CRWTeacherStudent  :294-604  a frozen pretrained CRWBase teacher beside the student; loss = alpha * walk loss +
                                 (1 - alpha) * soft cross-entropy of the teacher's chain products against the student's

Node formation stays as in model.CRW; everything from the node vectors on - both walks, the two losses and the gradient of
the student's node vectors - runs in libcrw_b200.so (ops.walk_teacher_student -> crw_walk_ts_fwd_bwd).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .model import CRW


class CRWBase(CRW):
    """teacherstudent.py:11-268: softmax instead of ZeroSoftmax, nn.Linear head WITH bias (checkpoints carry
    `selfsim_fc.0.bias`).  The biased head runs as the stock nn.Sequential."""

    def __init__(self, args, vis=None):
        super().__init__(args, vis)
        self.use_softmax = True

    def make_head(self, depth=1):
        head = []
        if depth >= 0:
            dims = [self.enc_hid_dim] + [self.enc_hid_dim] * depth + [128]
            for d1, d2 in zip(dims, dims[1:]):
                head += [nn.Linear(d1, d2), nn.ReLU()]
            head = head[:-1]
        return nn.Sequential(*head)

    def _head(self, pooled):
        return self.selfsim_fc(pooled)


class SoftCrossEntropyLoss(nn.modules.loss._Loss):
    """teacherstudent.py:270-292: -(target * log_softmax(input)).sum(dim), reduced by 'mean' | 'sum' | 'none'.  A plain
    module for callers that use it on its own; the walk computes the same quantity inside its loss kernel."""

    def __init__(self, size_average=None, reduce=None, reduction: str = "mean"):
        super().__init__(size_average, reduce, reduction)

    def forward(self, input, target, dim: int = -1):
        loss = (-target * F.log_softmax(input, dim=dim)).sum(dim)
        if self.reduction == "mean":
            return loss.mean()
        if self.reduction == "sum":
            return loss.sum()
        if self.reduction != "none":
            raise ValueError(self.reduction + " is not valid reduction")
        return loss


class CRWTeacherStudent(CRWBase):
    """teacherstudent.py:294-604.  `args.path_to_pretrained` is a checkpoint whose 'model' entry is a CRWBase state dict
    (:321-323); `args.alpha_teacher_student` in [0, 1] (:333-334).  A ready teacher module may be passed instead."""

    def __init__(self, args, vis=None, teacher: CRWBase = None):
        super().__init__(args, vis)
        if teacher is None:
            teacher = CRWBase(args)
            state = torch.load(args.path_to_pretrained, map_location="cpu")
            teacher.load_state_dict(state["model"])
        self.teacher = teacher.to(self.args.device)
        for param in self.teacher.parameters():
            param.requires_grad = False
        self.soft_xent = SoftCrossEntropyLoss()
        self.alpha = args.alpha_teacher_student
        assert 0 <= self.alpha <= 1, "alpha_teacher_student must be in the interval [0, 1]"

    def pixels_to_nodes_tchr(self, x):
        """teacherstudent.py:439-470: the teacher's node embeddings and maps, no gradient and NO feature dropout (:453-455)."""
        with torch.no_grad():
            return self.teacher.pixels_to_nodes(x, featdrop=False)

    def forward(self, x, just_feats=False, walk_uniforms=None):
        """x (B,T,N*3,H,W) -> (q, loss[1], diags) as teacherstudent.py:472-580; `walk_uniforms` optionally supplies the
        dropout draws (us12, us21p, ut12, ut21p), each (T-1,B,N,N)."""
        if not x.is_cuda:
            raise RuntimeError("CRWTeacherStudent.forward needs CUDA tensors on an sm_100 device: there is no CPU path")
        B, T, C, H, W = x.shape
        _N, C = C // 3, 3
        x = x.transpose(1, 2).reshape(B, _N, C, T, H, W)
        f, mm, _, _ = self._patch_nodes_prenorm(x)
        with torch.no_grad():
            ft = self.teacher._patch_nodes_prenorm(x, featdrop=False)[0]      # deterministic teacher (:453-455)
        qn, loss, xent, acc, ts_xent = ops.walk_teacher_student(f, ft, self.temperature, self.edgedrop_rate, self.alpha, flip=self.flip,
                                                                softmax=True, rng=self._rng_mode(f.device), uniforms=walk_uniforms)
        q = qn.permute(0, 3, 2, 1)                                               # (B, D, T, N)
        diags = dict()
        tag = "l" if self.flip else "r"
        for j in range(xent.shape[0]):
            diags["%s xent cyc %s%d" % (H, tag, j + 1)] = xent[j]
            diags["%s acc cyc %s%d" % (H, tag, j + 1)] = acc[j]
        return q, loss, diags
