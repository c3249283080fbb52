"""L1L2 regulariser tag (reference: utils/Regularizers.py).  As in the reference it is stored by
Dense (layers/FC.py:70) and never applied to the loss or the gradients."""


class Regularizer(object):
    pass


class L1L2(Regularizer):
    def __init__(self, l1=0., l2=0.):
        self.l1, self.l2 = l1, l2


def get_regularizer(regularizer):
    if regularizer is None or isinstance(regularizer, Regularizer):
        return regularizer
    if isinstance(regularizer, str):
        key = regularizer.lower()
        if key == 'l1':
            return L1L2(l1=0.01)
        if key == 'l2':
            return L1L2(l2=0.01)
        if key in ('l1l2', 'l1_l2'):
            return L1L2(0.01, 0.01)
    raise ValueError('unknown regularizer type!')
