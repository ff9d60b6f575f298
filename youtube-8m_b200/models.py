"""Contains the base class for models (the plugin contract of the reference, wh/models.py:17-21)."""


class BaseModel(object):
  """Inherit from this class when implementing new models."""

  def create_model(self, unused_model_input, **unused_params):
    raise NotImplementedError()
