from ...models._placeholder import placeholder

StableDiffusionSafetyChecker = placeholder("StableDiffusionSafetyChecker")
