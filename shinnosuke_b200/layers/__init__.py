from .Base import *          # noqa: F401,F403
from .Base import Layer, Variable, Input, Add
from .Activation import Activation
from .FC import Flatten, Dense
from .Convolution import Conv2D, ZeroPadding2D, MaxPooling2D
from .Normalization import BatchNormalization
from .Embeddings import Embedding
from .Recurrent import LSTM, LSTMCell, Recurrent

name = 'shinnosuke-b200-layers'
