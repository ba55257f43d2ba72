from .adn import ADN  # noqa: F401
from .convolutions import Convolution  # noqa: F401
