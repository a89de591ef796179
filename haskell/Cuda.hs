{-# LANGUAGE ForeignFunctionInterface #-}
-- | Graphics.Bling.Renderer.Cuda -- the `Renderer` instance a bling maintainer adds to drive libblingcu.so.
--
-- NOT COMPILED IN THIS REPOSITORY (no GHC in the build image, SURVEY.md F7). It is written against
-- include/blingcu.h and mirrors `prender` (Graphics/Bling/Rendering.hs:111-140): upload the flat scene once, then
-- loop { render one pass on the GPU(s); sum and read the film; wrap it into an `Image`; report PassDone } until the
-- reporter returns False. Every foreign call is `safe`: kernels run for a long time and must not block the RTS.
--
-- What this module needs from the rest of bling that does not exist today (INTEGRATION.md §2 lists the same three things):
--   * `RenderJob` gains ONE field, `jobSceneIR :: SceneIR` (Rendering.hs:37-41; `mkJob` gains the argument), filled by
--     IO/RenderJob.hs from the parser state;
--   * the parser state `PState` (IO/ParserCore.hs) gains ONE field, `sceneIR :: SceneIR`, appended to by the hooks of
--     haskell/SceneIR.hs wherever the parser builds a closure;
--   * `Graphics.Bling.Image` exports `imageFromRaw :: Int -> Int -> Filter -> V.Vector Float -> Image` (`Img` is not exported).
-- The executable stand-in for all of this is bling_b200/renderer.py (same loop, same calls, tested on the GPU).
module Graphics.Bling.Renderer.Cuda ( CudaRenderer, mkCudaRenderer ) where

import Control.Monad (forM, forM_, when, unless)
import Data.Word
import Data.Int
import Foreign
import Foreign.C.String
import Foreign.C.Types
import qualified Data.Vector.Storable as SV
import qualified Data.Vector.Storable.Mutable as SVM
import qualified Data.Vector.Unboxed as V
import qualified Text.PrettyPrint as PP

import Graphics.Bling.Image      -- needs one new export: imageFromRaw (see above)
import Graphics.Bling.Rendering  -- RenderJob with the new field jobSceneIR
import Graphics.Bling.Types
import Graphics.Bling.SceneIR    -- new module: the reified flat scene the parser emits (haskell/SceneIR.hs)

data Ctx  -- opaque blingcu_ctx

foreign import ccall safe "blingcu_create"        c_create       :: CInt -> Ptr (Ptr Ctx) -> IO CInt
foreign import ccall safe "blingcu_destroy"       c_destroy      :: Ptr Ctx -> IO ()
foreign import ccall safe "blingcu_last_error"    c_last_error   :: Ptr Ctx -> IO CString
foreign import ccall safe "blingcu_upload_scene"  c_upload_scene :: Ptr Ctx -> Ptr SceneIR -> IO CInt
foreign import ccall safe "blingcu_render_pass"   c_render_pass  :: Ptr Ctx -> Word32 -> Word64 -> IO CInt
foreign import ccall safe "blingcu_render_slice"  c_render_slice :: Ptr Ctx -> Word32 -> Word64 -> Word32 -> Word32 -> IO CInt
foreign import ccall safe "blingcu_read_film"     c_read_film    :: Ptr Ctx -> Ptr CFloat -> IO CInt
foreign import ccall safe "blingcu_trace_nearest" c_trace_nearest :: Ptr Ctx -> Ptr CFloat -> CSize -> Ptr CFloat -> IO CInt
foreign import ccall safe "blingcu_get_stats"     c_get_stats    :: Ptr Ctx -> Ptr Word64 -> IO CInt
-- parity hook: a texture-table entry at explicit (dgP, (dgU, dgV)) points, to hold against the Haskell `Texture a` closure
foreign import ccall safe "blingcu_eval_texture"  c_eval_texture :: Ptr Ctx -> Int32 -> Ptr CFloat -> Ptr CFloat -> CSize -> Ptr CFloat -> IO CInt
-- multi-GPU (include/blingcu.h "multi-GPU"): one process, one context per device; the library owns NCCL
foreign import ccall safe "blingcu_comm_init_all"     c_comm_init_all     :: Ptr (Ptr Ctx) -> CInt -> IO CInt
foreign import ccall safe "blingcu_reduce_film_group" c_reduce_film_group :: Ptr (Ptr Ctx) -> CInt -> CInt -> IO CInt
foreign import ccall safe "blingcu_read_film_sum"     c_read_film_sum     :: Ptr Ctx -> Ptr CFloat -> IO CInt
-- the light tracer (Renderer/LightTracer.hs): photons [first, first + n) of a pass into the splat buffer, [H][W]{X, Y, Z}
foreign import ccall safe "blingcu_light_trace"       c_light_trace       :: Ptr Ctx -> Word32 -> Word64 -> Word64 -> Word32 -> IO CInt
foreign import ccall safe "blingcu_read_splat"        c_read_splat        :: Ptr Ctx -> Ptr CFloat -> IO CInt

-- | `renderer { cuda devices 0 1 2 3 seed 42 }` in a .bling file (IO/RendererParser.hs:26-51 gains one case)
data CudaRenderer = CR { crDevices :: [Int], crSeed :: Word64 }

mkCudaRenderer :: [Int] -> Word64 -> CudaRenderer
mkCudaRenderer = CR

instance Printable CudaRenderer where
   prettyPrint (CR ds _) = PP.text "cuda sampler renderer on devices" PP.<+> PP.hsep (map PP.int ds)

check :: Ptr Ctx -> CInt -> IO ()
check ctx rc = unless (rc == 0) $ do
   msg <- c_last_error ctx >>= peekCString
   ioError $ userError $ "blingcu error " ++ show rc ++ ": " ++ msg

createOn :: Int -> IO (Ptr Ctx)
createOn dev = alloca $ \pctx -> do
   rc <- c_create (fromIntegral dev) pctx
   when (rc /= 0) $ do
      msg <- c_last_error nullPtr >>= peekCString
      ioError $ userError $ "blingcu_create: " ++ msg      -- no CPU fallback: the caller picks another renderer
   peek pctx

-- | sample indices [s0, s1) of one pass owned by device r of n (bling_b200/renderer.py::shard_range)
shardRange :: Int -> Int -> Int -> (Word32, Word32)
shardRange spp r n = (fromIntegral ((spp * r) `div` n), fromIntegral ((spp * (r + 1)) `div` n))

instance Renderer CudaRenderer where
   render (CR devs seed) job report = do
      ctxs <- forM devs createOn
      let (w, h) = jobImageSize job
          ir = jobSceneIR job                                 -- NEW RenderJob field
          n = length ctxs
          (_, nu, nv) = irSampler ir
          spp = fromIntegral nu * fromIntegral nv :: Int
      withArrayLen ctxs $ \_ pctxs -> do
         when (n > 1) $ c_comm_init_all pctxs (fromIntegral n) >>= check (head ctxs)
         -- the flat scene was recorded by the parser while it built the closures (Primitive.hs:21-27 cannot be
         -- flattened afterwards); withSceneIR marshals it into a blingcu_scene for the duration of the call.
         -- Every device holds a full replica (SURVEY.md §8e).
         withSceneIR ir $ \pir -> forM_ ctxs $ \c -> c_upload_scene c pir >>= check c
         _ <- report Started
         let pass p = do
               -- render calls only enqueue work, so one host thread keeps every device busy
               forM_ (zip [0 ..] ctxs) $ \(r, c) -> do
                  let (s0, s1) = shardRange spp r n
                  when (s1 > s0) $ c_render_slice c (fromIntegral p) seed s0 s1 >>= check c
               mv <- SVM.new (w * h * 4)                         -- [H][W]{weight, X*w, Y*w, Z*w} == Img._imgP
               if n > 1
                  then do
                     c_reduce_film_group pctxs (fromIntegral n) 0 >>= check (head ctxs)   -- ncclReduce onto device 0
                     SVM.unsafeWith mv $ \ptr -> c_read_film_sum (head ctxs) (castPtr ptr) >>= check (head ctxs)
                  else SVM.unsafeWith mv $ \ptr -> c_read_film (head ctxs) (castPtr ptr) >>= check (head ctxs)
               film <- SV.unsafeFreeze mv
               let img = imageFromRaw w h (jobPixelFilter job) (V.convert film)
               cont <- report (PassDone p img 1)
               when cont $ pass (p + 1)
         pass (1 :: Int)
      forM_ ctxs c_destroy

-- | `renderer { cudaLight passPhotons n }`: Renderer/LightTracer.hs:39-51 on the device. The splat buffer comes back as
-- [H][W]{X, Y, Z} == Img._imgS (Image.hs:64-70), the film stays empty, and the reporter gets the same
-- `PassDone n img (1 / (n * ppp))` the CPU light tracer sends. The scene IR's camera must carry world2raster and the
-- pixel area (Camera.hs:78-103 sampleCam; `irCamera` fills them from the parser's transform).
data CudaLightTracer = CLT { cltDevice :: Int, cltPassPhotons :: Int, cltSeed :: Word64 }

instance Printable CudaLightTracer where
   prettyPrint (CLT d n _) = PP.vcat [PP.text "CUDA light tracer on device" PP.<+> PP.int d, PP.int n PP.<+> PP.text "photons per pass"]

instance Renderer CudaLightTracer where
   render (CLT dev ppp seed) job report = do
      c <- createOn dev
      let (w, h) = jobImageSize job
      withSceneIR (jobSceneIR job) $ \pir -> c_upload_scene c pir >>= check c
      _ <- report Started
      let pass p = do
            c_light_trace c (fromIntegral p) seed 0 (fromIntegral ppp) >>= check c
            ms <- SVM.new (w * h * 3)
            SVM.unsafeWith ms $ \ptr -> c_read_splat c (castPtr ptr) >>= check c
            splats <- SV.unsafeFreeze ms
            let img = imageFromSplats w h (jobPixelFilter job) (V.convert splats)      -- an Image with empty _imgP and these _imgS
            cont <- report (PassDone p img (1 / fromIntegral (p * ppp)))
            when cont $ pass (p + 1)
      pass (1 :: Int)
      c_destroy c
