-- | Graphics.Bling.SceneIR -- the reified, flat scene the parser records next to the closures it builds
-- (INTEGRATION.md §2.1), and its marshalling into a `blingcu_scene` (include/blingcu.h).
--
-- NOT COMPILED IN THIS REPOSITORY (no GHC in the build image, SURVEY.md F7). The record layouts come from
-- Graphics.Bling.Renderer.Cuda.Layout (haskell/Layout.hs), which `tools/abi_layout.py --haskell` generates from the header with
-- the C compiler's sizeof / offsetof; tests/test_host_and_emu.py::test_ctypes_mirror_matches_the_header_layout keeps that file
-- and the Python mirror in step with the header. bling_b200/host/loader.py is the executable stand-in for what the parser
-- hooks below record (same field meanings, same orderings).
module Graphics.Bling.SceneIR
   ( SceneIR(..), emptySceneIR
   , TextureIR(..), MaterialIR(..), ShapeIR(..), LightIR(..), MeshIR(..)
   , addTexture, addMaterial, addShape, addMesh, addLight
   , withSceneIR
   ) where

import Control.Monad (forM_, zipWithM_)
import Data.Int
import Data.Word
import Foreign
import Foreign.C.Types
import qualified Data.Vector.Storable as SV

import Graphics.Bling.Renderer.Cuda.Layout

-- | one `blingcu_texture`: kind (BLINGCU_TEX_* / BLINGCU_STEX_*), two children, aux, eight floats, sixteen floats.
--   Where the parser builds a closure (IO/MaterialParser.hs:113-232) it also appends one of these and remembers the index.
data TextureIR = TextureIR
   { texKind :: !Int32, texChild :: !(Int32, Int32), texAux :: !Int32
   , texF :: ![Float]      -- ^ up to 8: line width, uv mapping, scale / offset, gradient position, fbm omega ...
   , texS :: ![Float] }    -- ^ 16: a Spectrum (constant, gradient step), or a 2-D / 3-D mapping (see the header)

-- | one `blingcu_material`: texture indices, constant scalars, and 1-based scalar-texture references (0 = constant)
data MaterialIR = MaterialIR
   { matKind :: !Int32, matTex :: ![Int32], matF :: ![Float], matFTex :: ![Int32], matBump :: !Int32 }

-- | one analytic shape wrapped by mkGeom (Primitive/Geometry.hs:14-36); matrices row-major as Transform.hs:40-42
data ShapeIR = ShapeIR
   { shpKind :: !Int32, shpMaterial :: !Int32, shpLight :: !Int32, shpPrimId :: !Int32
   , shpParams :: ![Float], shpO2W :: ![Float], shpW2O :: ![Float] }

-- | world-space triangles of one mesh, in the fan order of TriangleMesh.hs:23-29
data MeshIR = MeshIR
   { meshVerts :: !(SV.Vector Float)            -- ^ 9 per triangle
   , meshUVs :: !(SV.Vector Float)              -- ^ 6 per triangle (default 0,0,1,0,1,1: TriangleMesh.hs:119-120)
   , meshNormals :: !(Maybe (SV.Vector Float))  -- ^ 9 per triangle; when some meshes of a scene have normals and others do not,
                                                --   the marshaller writes nine zeros per triangle for a `Nothing` (blingcu.h: flat shaded)
   , meshMaterial :: !Int32, meshFirstPrimId :: !Int32 }

data LightIR = LightIR { lgtKind :: !Int32, lgtShape :: !Int32, lgtEnv :: !Int32, lgtV :: ![Float], lgtS :: ![Float] }

data SceneIR = SceneIR
   { irTextures :: [TextureIR], irMaterials :: [MaterialIR], irShapes :: [ShapeIR], irMeshes :: [MeshIR]
   , irLights :: [LightIR]                      -- ^ sceneLights order (Scene.hs:38-42)
   , irImageSize :: (Int, Int), irFilterSize :: (Float, Float), irFilterTable :: [Float]   -- mkTableFilter (Image.hs:40-61)
   , irSampler :: (Int32, Int32, Int32)         -- ^ kind, nu, nv
   , irIntegrator :: (Int32, Int32, Int32)      -- ^ BLINGCU_INTEGRATOR_*, maxDepth, sampleDepth
   , irCamera :: [Word8]                        -- ^ a marshalled blingcu_camera (CameraParser.hs:30-38)
   , irCie :: ([Float], [Float], [Float], Float), irIllumBasis :: [[Float]], irReflBasis :: [[Float]] }

emptySceneIR :: SceneIR
emptySceneIR = SceneIR [] [] [] [] [] (0, 0) (0.5, 0.5) (replicate 256 1) (0, 2, 2) (0, 7, 3) [] ([], [], [], 1) [] []

-- the parser state (IO/ParserCore.hs:45-58) threads a SceneIR; each hook returns the index the caller stores in its records
addTexture :: TextureIR -> SceneIR -> (Int32, SceneIR)
addTexture t ir = (fromIntegral (length (irTextures ir)), ir { irTextures = irTextures ir ++ [t] })

addMaterial :: MaterialIR -> SceneIR -> (Int32, SceneIR)
addMaterial m ir = (fromIntegral (length (irMaterials ir)), ir { irMaterials = irMaterials ir ++ [m] })

-- | primitives are PREPENDED block by block, like `p ++ prims s` in IO/RenderJob.hs:49-52: the prim id of a record is its
--   position in the list handed to mkScene, so ids are assigned in `withSceneIR`, not here
addShape :: ShapeIR -> SceneIR -> SceneIR
addShape s ir = ir { irShapes = s : irShapes ir }

addMesh :: MeshIR -> SceneIR -> SceneIR
addMesh m ir = ir { irMeshes = m : irMeshes ir }

addLight :: LightIR -> SceneIR -> SceneIR
addLight l ir = ir { irLights = l : irLights ir }   -- ls : lights s (IO/LightParser.hs)

pokeFloats :: Ptr a -> Int -> [Float] -> IO ()
pokeFloats p off xs = zipWithM_ (\i x -> pokeByteOff p (off + 4 * i) (realToFrac x :: CFloat)) [0 ..] xs

pokeInts :: Ptr a -> Int -> [Int32] -> IO ()
pokeInts p off xs = zipWithM_ (\i x -> pokeByteOff p (off + 4 * i) x) [0 ..] xs

pokeTexture :: Ptr a -> TextureIR -> IO ()
pokeTexture p (TextureIR k (c0, c1) aux f s) = do
   fillBytes p 0 sizeOfTexture
   pokeByteOff p offTextureKind k
   pokeInts p offTextureChild [c0, c1]
   pokeByteOff p offTextureAux aux
   pokeFloats p offTextureF f
   pokeFloats p (offTextureS + offSpectrumV) s

pokeMaterial :: Ptr a -> MaterialIR -> IO ()
pokeMaterial p (MaterialIR k tex f ftex bump) = do
   fillBytes p 0 sizeOfMaterial
   pokeByteOff p offMaterialKind k
   pokeInts p offMaterialTex (take 3 (tex ++ repeat (-1)))
   pokeByteOff p offMaterialTex3 (if length tex > 3 then tex !! 3 else -1 :: Int32)
   pokeFloats p offMaterialF f
   pokeInts p offMaterialFtex ftex
   pokeByteOff p offMaterialBump bump

pokeShape :: Ptr a -> ShapeIR -> IO ()
pokeShape p (ShapeIR k m l pid ps o2w w2o) = do
   fillBytes p 0 sizeOfShape
   pokeInts p offShapeKind [k, m, l, pid]        -- kind, material, light, prim_id are consecutive int32
   pokeFloats p offShapeP ps
   pokeFloats p offShapeO2w o2w
   pokeFloats p offShapeW2o w2o

pokeLight :: Ptr a -> LightIR -> IO ()
pokeLight p (LightIR k sh env v s) = do
   fillBytes p 0 sizeOfLight
   pokeInts p offLightKind [k, sh, env]
   pokeFloats p offLightV v
   pokeFloats p (offLightS + offSpectrumV) s

-- | a C array of fixed-size records, alive for the duration of the action
withRecords :: Int -> (Ptr () -> a -> IO ()) -> [a] -> (Ptr () -> IO b) -> IO b
withRecords size pokeOne xs act = allocaBytes (max 1 (size * length xs)) $ \p -> do
   forM_ (zip [0 ..] xs) $ \(i, x) -> pokeOne (p `plusPtr` (i * size)) x
   act p

-- | marshals the IR into a `blingcu_scene` that lives for the duration of the action (blingcu_upload_scene copies everything).
--   Environment maps (blingcu_envmap with the Dist2D arrays of Montecarlo.hs:34-104) and images follow the same pattern and are
--   left out of this sketch.
withSceneIR :: SceneIR -> (Ptr SceneIR -> IO a) -> IO a
withSceneIR ir act =
   let tris = SV.concat (map meshVerts (irMeshes ir)); uvs = SV.concat (map meshUVs (irMeshes ir))
       ntri = SV.length tris `div` 9
       triMat = SV.fromList (concatMap (\m -> replicate (SV.length (meshVerts m) `div` 9) (meshMaterial m)) (irMeshes ir))
       (w, h) = irImageSize ir; (fw, fh) = irFilterSize ir
       (sk, nu, nv) = irSampler ir; (ik, md, sd) = irIntegrator ir
       (cx, cy, cz, ysum) = irCie ir
   in allocaBytes sizeOfScene $ \p ->
      SV.unsafeWith tris $ \pv -> SV.unsafeWith uvs $ \pu -> SV.unsafeWith triMat $ \pm ->
      withRecords sizeOfShape pokeShape (irShapes ir) $ \pShapes ->
      withRecords sizeOfMaterial pokeMaterial (irMaterials ir) $ \pMats ->
      withRecords sizeOfTexture pokeTexture (irTextures ir) $ \pTex ->
      withRecords sizeOfLight pokeLight (irLights ir) $ \pLights -> do
         fillBytes p 0 sizeOfScene
         pokeByteOff p offSceneNTriangles (fromIntegral ntri :: Word64)
         pokeByteOff p offSceneTriVerts pv; pokeByteOff p offSceneTriUvs pu; pokeByteOff p offSceneTriMaterial pm
         pokeByteOff p offSceneNShapes (fromIntegral (length (irShapes ir)) :: Word32); pokeByteOff p offSceneShapes pShapes
         pokeByteOff p offSceneNMaterials (fromIntegral (length (irMaterials ir)) :: Word32); pokeByteOff p offSceneMaterials pMats
         pokeByteOff p offSceneNTextures (fromIntegral (length (irTextures ir)) :: Word32); pokeByteOff p offSceneTextures pTex
         pokeByteOff p offSceneNLights (fromIntegral (length (irLights ir)) :: Word32); pokeByteOff p offSceneLights pLights
         zipWithM_ (\i b -> pokeByteOff p (offSceneCamera + i) b) [0 ..] (irCamera ir)
         pokeInts p offSceneWidth [fromIntegral w, fromIntegral h]
         pokeFloats p offSceneFilterW [fw, fh]
         pokeFloats p offSceneFilterTable (irFilterTable ir)
         pokeInts p offSceneSamplerKind [sk, nu, nv, md, sd]     -- sampler_kind, nu, nv, max_depth, sample_depth are consecutive
         pokeFloats p (offSceneCieX + offSpectrumV) cx; pokeFloats p (offSceneCieY + offSpectrumV) cy
         pokeFloats p (offSceneCieZ + offSpectrumV) cz; pokeFloats p offSceneCieYSum [ysum]
         zipWithM_ (\i b -> pokeFloats p (offSceneIllumBasis + i * sizeOfSpectrum) b) [0 ..] (irIllumBasis ir)
         zipWithM_ (\i b -> pokeFloats p (offSceneReflBasis + i * sizeOfSpectrum) b) [0 ..] (irReflBasis ir)
         pokeByteOff p offSceneIntegratorKind ik
         act (castPtr p)
